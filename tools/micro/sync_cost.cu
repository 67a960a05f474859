// Microbenchmark: cost of one link of a dependent chain on B200
//   (a) grid-wide barrier inside a persistent kernel (148 CTAs, atomic arrive + acquire spin)
//   (b) kernel boundary with programmatic dependent launch inside a CUDA graph
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_cost sync_cost.cu && ./sync_cost
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
    }
    __syncthreads();
}
// work = 0: barrier only; work = 1: every CTA writes 448 B before and reads a 57 KB vector (all CTAs' slices) after
__global__ void __launch_bounds__(256, 1) persistent_k(unsigned* ctr, int iters, int work, float* buf, float* sink) {
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (work) {
            float* dst = buf + (size_t)(it & 1) * 16384;
            if (threadIdx.x < 112) dst[blockIdx.x * 112 + threadIdx.x] = acc + it;      // 148 * 112 = 16576 floats ~ 8 x 3584 bf16
        }
        grid_barrier(ctr, (unsigned)(it + 1) * gridDim.x);
        if (work) {
            const float4* src = reinterpret_cast<const float4*>(buf + (size_t)(it & 1) * 16384);
            for (int i = threadIdx.x; i < 4096; i += 256) {
                float4 v;
                asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i));
                acc += v.x + v.y + v.z + v.w;
            }
        }
    }
    if (acc == 123.f) *sink = acc;
}
__global__ void link_k(float* buf, int n) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x < 112) buf[blockIdx.x * 112 + threadIdx.x] += 1.f;
}
int main() {
    unsigned* ctr; float *buf, *sink;
    cudaMalloc(&ctr, 4); cudaMalloc(&buf, 1 << 20); cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, 1 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int work = 0; work < 2; ++work) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(ctr, 0, 4);
            const int iters = 2000;
            cudaEventRecord(e0);
            persistent_k<<<148, 256>>>(ctr, iters, work, buf, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("grid barrier (148 CTAs, work=%d): %.3f us per barrier\n", work, ms * 1e3 / iters);
        }
    }
    cudaStream_t st; cudaStreamCreate(&st);
    for (int pdl = 0; pdl < 2; ++pdl) for (int grid : {8, 148}) {
        const int n = 500;
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        for (int i = 0; i < n; ++i) {
            cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
            cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = a; cfg.numAttrs = pdl;
            cudaLaunchKernelEx(&cfg, link_k, buf, n);
        }
        cudaStreamEndCapture(st, &g);
        cudaGraphInstantiate(&ge, g, 0);
        cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
        cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("graph chain of %d kernels (grid %3d, pdl=%d): %.3f us per link\n", n, grid, pdl, ms * 1e3 / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
