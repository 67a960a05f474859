// Microbenchmark: does an L2 prefetch make a later streaming read of the same range faster?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2pf_test l2pf_test.cu && ./l2pf_test
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void flush_k(uint4* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(i, 1, 2, 3);
}
__global__ void read_k(const uint4* p, size_t n, uint32_t* out) {
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345) *out = acc;
}
__global__ void pf_bulk_k(const char* p, size_t bytes, uint32_t chunk) {
    size_t nchunks = bytes / chunk;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nchunks; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + i * chunk), "r"(chunk) : "memory");
}
__global__ void pf_line_k(const char* p, size_t bytes) {
    size_t n = bytes / 128;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i * 128) : "memory");
}
int main() {
    const size_t big = 1ull << 30, tgt = 32ull << 20;
    char *f, *t; uint32_t* out;
    cudaMalloc(&f, big); cudaMalloc(&t, tgt); cudaMalloc(&out, 4);
    cudaMemset(t, 1, tgt);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto flush = [&] { flush_k<<<1184, 256>>>((uint4*)f, big / 16); cudaDeviceSynchronize(); };
    auto timed_read = [&](const char* name) {
        cudaEventRecord(e0);
        read_k<<<592, 256>>>((const uint4*)t, tgt / 16, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s read of 32 MiB: %7.2f us  (%.0f GB/s)\n", name, ms * 1e3, tgt / ms / 1e6);
    };
    for (int rep = 0; rep < 2; ++rep) {
        flush(); timed_read("cold (after 1 GiB flush)");
        timed_read("warm (just read)");
        for (uint32_t chunk : {128u, 1024u, 2560u, 7168u, 16384u}) {
            flush();
            pf_bulk_k<<<256, 256>>>(t, tgt, chunk); cudaDeviceSynchronize();
            char nm[64]; snprintf(nm, 64, "after cp.async.bulk.prefetch.L2 chunk=%u", chunk);
            timed_read(nm);
        }
        flush(); pf_line_k<<<256, 256>>>(t, tgt); cudaDeviceSynchronize(); timed_read("after prefetch.global.L2 per 128 B");
        // prefetch timing itself
        flush();
        cudaEventRecord(e0); pf_bulk_k<<<256, 256>>>(t, tgt, 2560); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); printf("bulk prefetch kernel itself: %.2f us\n", ms * 1e3);
        flush();
        cudaEventRecord(e0); pf_line_k<<<256, 256>>>(t, tgt); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("line prefetch kernel itself: %.2f us\n", ms * 1e3);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
