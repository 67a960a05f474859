// Read-only HBM streaming ceiling on this GPU: (a) coalesced 16-byte loads, (b) 1-D bulk copies (cp.async.bulk) into a shared-memory
// ring, (c) device-to-device memcpy (read + write, the MEASURED_PEAKS.json definition).  nvcc -O3 -arch=sm_100a read_bw.cu -o read_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void ldg_stream(const uint4* __restrict__ p, size_t n, unsigned* sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n; i += 8 * stride) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int kStages, int kChunk>
__global__ void __launch_bounds__(128, 1) bulk_stream(const uint8_t* __restrict__ p, size_t bytes, unsigned* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[kStages];
    const size_t chunks = bytes / kChunk;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    size_t c = blockIdx.x;
    int issued = 0, done = 0;
    unsigned phase_bits = 0;
    auto issue = [&](size_t chunk, int s) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(kChunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + (size_t)s * kChunk)), "l"(p + chunk * kChunk), "r"(kChunk), "r"(smem_u32(&bar[s])) : "memory");
    };
    for (; issued < kStages && c < chunks; ++issued, c += gridDim.x) issue(c, issued);
    while (done < issued) {
        const int s = done % kStages;
        const unsigned ph = (phase_bits >> s) & 1u;
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(ph) : "memory");
        }
        phase_bits ^= 1u << s;
        ++done;
        if (c < chunks) { issue(c, s); ++issued; c += gridDim.x; }
    }
    if (smem[5] == 77 && smem[kChunk + 1] == 3) *sink = 1;
}

int main() {
    const size_t bytes = (size_t)6 << 30;
    uint8_t *a, *b;
    unsigned* sink;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    auto report = [&](const char* name, double moved) { printf("%-44s %8.1f GB/s\n", name, moved / (ms / 1e3) / 1e9); };
    for (int rep = 0; rep < 2; ++rep) {
        for (int mult : {2, 4, 8, 16}) {
            for (int thr : {256, 512}) {
                ldg_stream<<<148 * mult, thr>>>((const uint4*)a, bytes / 16, sink);
                cudaEventRecord(e0); ldg_stream<<<148 * mult, thr>>>((const uint4*)b, bytes / 16, sink); cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                char nm[96]; snprintf(nm, 96, "ldg 16B x8 unroll, %d CTAs/SM x %d thr", mult, thr);
                if (rep) report(nm, (double)bytes);
            }
        }
#define BULK(ST, CH) { CK(cudaFuncSetAttribute(bulk_stream<ST, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * CH)); \
            bulk_stream<ST, CH><<<148, 128, ST * CH>>>(a, bytes, sink); cudaEventRecord(e0); bulk_stream<ST, CH><<<148, 128, ST * CH>>>(b, bytes, sink); \
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); \
            char nm[96]; snprintf(nm, 96, "bulk 1-D %d x %d KB ring, 1 CTA/SM", ST, CH / 1024); if (rep) report(nm, (double)bytes); }
        BULK(4, 16384) BULK(8, 16384) BULK(12, 16384) BULK(6, 32768) BULK(24, 8192)
        cudaEventRecord(e0); CK(cudaMemcpyAsync(b, a, bytes / 2, cudaMemcpyDeviceToDevice)); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        if (rep) report("cudaMemcpy D2D (read + write bytes)", (double)bytes);
    }
    return 0;
}
