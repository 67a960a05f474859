// Shared device helpers: bf16 rounding points, warp/block reductions, vector access.
#pragma once
#include <cstdint>
#include <cstdio>
#include <utility>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace umv {

using bf16 = __nv_bfloat16;

// Last error text of the calling thread (C-ABI: umv_last_error()).
void set_error(const char* fmt, ...);
const char* get_error();

#define UMV_CUDA_OK(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            umv::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return UMV_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define UMV_REQUIRE(cond, code, ...)       \
    do {                                   \
        if (!(cond)) {                     \
            umv::set_error(__VA_ARGS__);   \
            return (code);                 \
        }                                  \
    } while (0)

// ---- bf16 rounding points (round-to-nearest-even, what torch's .to(bfloat16) does) ----
__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float b2f(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ bf16 f2b(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

// Packed bf16 arithmetic with ONE rounding per lane.  For bf16 operands the fp32 product (16 significant bits) and the fp32 sum
// relevant here are exact before rounding, so mul.rn.bf16x2 == bf16(float(a) * float(b)) and add.rn.bf16x2 == bf16(float(a) + float(b)):
// the reference's "compute in fp32, store bf16" steps on bf16 inputs at half the instructions (no unpack, no separate convert).
__device__ __forceinline__ uint32_t bmul2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t badd2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// torch's silu on a bf16 tensor: fp32 x / (1 + exp(-x)), one rounding.
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
// torch gelu(approximate="tanh") opmath: 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
__device__ __forceinline__ float gelu_tanh_f(float x) {
    const float kBeta = 0.7978845608028654f * 1.0f;
    const float kKappa = 0.044715f;
    float inner = kBeta * (x + kKappa * x * x * x);
    return 0.5f * x * (1.0f + tanhf(inner));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Block-wide sum; `red` is >= 32 floats of shared memory.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : 0.0f;
    t = warp_sum(t);
    return t;
}

// Programmatic dependent launch: every kernel of the engine starts with pdl_launch_dependents() +
// pdl_wait() (ptx.cuh), so back-to-back launches overlap their launch latency / prologue -- and, for the
// weight-major linear, the first weight tiles -- with the tail of the previous kernel.  UMV_PDL=0 disables.
extern bool g_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- timeline trace (diagnostic; umv_trace_begin / umv_trace_read): with a non-null slot a kernel stamps %globaltimer
// when its first CTA starts, when that CTA passes griddepcontrol.wait, when it ends, and the latest end over all CTAs.
struct TraceSlot {
    unsigned long long t_start, t_wait, t_end_first, t_end_last;
    unsigned long long dbg[8];      // kernel-specific intermediate stamps of the first CTA
};
bool trace_active();                               // host: a trace is being recorded (slot pointers get baked into launches)
TraceSlot* trace_next(const char* kernel_name);     // host: next slot of the active trace, or nullptr (tracing off)
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool trace_lead() { return threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0; }
__device__ __forceinline__ void trace_start(TraceSlot* s) { if (s && trace_lead()) s->t_start = gtime(); }
__device__ __forceinline__ void trace_dbg(TraceSlot* s, int i) { if (s && trace_lead()) s->dbg[i] = gtime(); }
__device__ __forceinline__ void trace_wait(TraceSlot* s) { if (s && trace_lead()) s->t_wait = gtime(); }
// latest time over ALL CTAs at which thread 0 passes this point (stamps grow monotonically, so after graph replays the slot holds the
// last replay's value)
__device__ __forceinline__ void trace_dbg_max(TraceSlot* s, int i) { if (s && threadIdx.x == 0) atomicMax(&s->dbg[i], gtime()); }
// kSync = false for kernels whose threads may have returned early (thread 0's own end is stamped)
template <bool kSync = true>
__device__ __forceinline__ void trace_end(TraceSlot* s) {
    if (!s) return;
    if (kSync) __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t = gtime();
        atomicMax(&s->t_end_last, t);
        if (trace_lead()) s->t_end_first = t;
    }
}
#endif

struct alignas(16) U4 {
    uint32_t x, y, z, w;
};
__device__ __forceinline__ U4 ldg16(const void* p) { return *reinterpret_cast<const U4*>(p); }
__device__ __forceinline__ void stg16(void* p, const U4& v) { *reinterpret_cast<U4*>(p) = v; }
__device__ __forceinline__ U4 ldg16_stream(const void* p) {
    U4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace umv
