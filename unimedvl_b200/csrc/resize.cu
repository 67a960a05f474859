// Device-side image resize in front of the ViT / VAE transforms.
//
// Reference call site: MaxLongEdgeMinShortEdgeResize.forward -> torchvision F.resize(PIL image, BICUBIC, antialias=True)
// (codes/data/transforms.py:60-87), i.e. PIL.Image.resize((w, h), BICUBIC) -- Pillow's 8-bit resampler: per-axis tap
// windows and weights in double precision (support widened by the scale when shrinking), weights rounded to fixed point
// with 22 fractional bits, a horizontal pass into an 8-bit intermediate image, then a vertical pass.  Everything after the
// weight table is integer arithmetic, so the device result is bit-identical to Pillow's (oracle/resize.py, tests/test_resize.py).
//
// HBM-bound byte work: one thread per output pixel (3 channels), taps read through L1/L2 (neighbouring outputs share them),
// coalesced along x in both passes.  A 448 x 448 target is 0.6 MB -- the kernel exists to keep the image on the device
// between the uint8 upload and the patch matrix, not because it is hot.
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/umv.h"
#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace umv {

#define UMV_LAUNCH_CHECK(name)                                                        \
    do {                                                                              \
        ++g_launches;                                                                 \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            set_error("%s launch failed: %s", name, cudaGetErrorString(_e));          \
            return UMV_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

inline double bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

int axis_ksize(int in_size, int out_size) {
    double scale = (double)((float)in_size - 0.0f) / out_size;
    if (scale < 1.0) scale = 1.0;
    return (int)std::ceil(2.0 * scale) * 2 + 1;
}

// bounds[out][2] = (first tap, tap count), kk[out][ksize] fixed-point weights
void axis_coefficients(int in_size, int out_size, int32_t* bounds, int32_t* kk, int ksize) {
    const double scale = (double)((float)in_size - 0.0f) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 2.0 * filterscale;
    const double ss = 1.0 / filterscale;
    std::vector<double> w(ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            w[x] = bicubic((x + xmin - center + 0.5) * ss);
            ww += w[x];
        }
        int32_t* k = kk + (size_t)xx * ksize;
        for (int x = 0; x < xmax; ++x) {
            const double v = ww != 0.0 ? w[x] / ww : w[x];
            k[x] = v < 0 ? (int32_t)(-0.5 + v * (1 << kPrecisionBits)) : (int32_t)(0.5 + v * (1 << kPrecisionBits));
        }
        for (int x = xmax; x < ksize; ++x) k[x] = 0;
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= kPrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// one pass along `axis`: out[y][x][c] = clip8(2^21 + sum_i in[tap_i][c] * k[i]).  HORIZ: taps run along x, else along y.
template <bool HORIZ>
__global__ void __launch_bounds__(128) resample_kernel(const uint8_t* __restrict__ in, int in_w, int out_h, int out_w,
                                                        const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk,
                                                        int ksize, uint8_t* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= out_w) return;
    const int o = HORIZ ? x : y;
    const int first = bounds[2 * o], n = bounds[2 * o + 1];
    const int32_t* k = kk + (size_t)o * ksize;
    const uint8_t* p = HORIZ ? in + ((size_t)y * in_w + first) * 3 : in + ((size_t)first * in_w + x) * 3;
    const size_t step = HORIZ ? 3 : (size_t)in_w * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int i = 0; i < n; ++i, p += step) {
        const int kv = k[i];
        s0 += (int)p[0] * kv;
        s1 += (int)p[1] * kv;
        s2 += (int)p[2] * kv;
    }
    uint8_t* d = out + ((size_t)y * out_w + x) * 3;
    d[0] = clip8(s0);
    d[1] = clip8(s1);
    d[2] = clip8(s2);
}

struct Layout {
    size_t bh, kh, bv, kv, tmp, total;
    int ksh, ksv;
};
Layout layout(int in_h, int in_w, int out_h, int out_w) {
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    Layout l{};
    l.ksh = axis_ksize(in_w, out_w);
    l.ksv = axis_ksize(in_h, out_h);
    size_t off = 0;
    l.bh = off; off = up(off + (size_t)out_w * 2 * 4);
    l.kh = off; off = up(off + (size_t)out_w * l.ksh * 4);
    l.bv = off; off = up(off + (size_t)out_h * 2 * 4);
    l.kv = off; off = up(off + (size_t)out_h * l.ksv * 4);
    l.tmp = off; off = up(off + (size_t)in_h * out_w * 3);
    l.total = off;
    return l;
}

}  // namespace
}  // namespace umv

using namespace umv;

extern "C" {

int umv_resize_coefficients(int32_t in_size, int32_t out_size, int32_t* bounds, int32_t* kk, int32_t* ksize) {
    UMV_REQUIRE(in_size > 0 && out_size > 0 && ksize, UMV_ERR_INVALID, "umv_resize_coefficients: bad argument");
    const int ks = axis_ksize(in_size, out_size);
    *ksize = ks;
    if (bounds && kk) axis_coefficients(in_size, out_size, bounds, kk, ks);
    return UMV_OK;
}

int64_t umv_resize_workspace_bytes(int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w) {
    if (in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0) return 0;
    return (int64_t)layout(in_h, in_w, out_h, out_w).total;
}

int umv_resize_bicubic_u8(const uint8_t* src, int32_t in_h, int32_t in_w, uint8_t* dst, int32_t out_h, int32_t out_w,
                          void* workspace, int64_t workspace_bytes, void* stream) {
    UMV_REQUIRE(src && dst && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, UMV_ERR_INVALID, "umv_resize_bicubic_u8: null/empty argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_h == out_h && in_w == out_w) {              // PIL.Image.resize returns a copy
        UMV_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)in_h * in_w * 3, cudaMemcpyDeviceToDevice, st));
        return UMV_OK;
    }
    const Layout l = layout(in_h, in_w, out_h, out_w);
    UMV_REQUIRE(workspace && workspace_bytes >= (int64_t)l.total, UMV_ERR_NOMEM,
                "umv_resize_bicubic_u8: workspace of %lld bytes, %zu needed (umv_resize_workspace_bytes)", (long long)workspace_bytes, l.total);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    const bool horiz = in_w != out_w, vert = in_h != out_h;
    std::vector<int32_t> host((l.tmp + 3) / 4, 0);
    uint8_t* hb = reinterpret_cast<uint8_t*>(host.data());
    if (horiz) axis_coefficients(in_w, out_w, reinterpret_cast<int32_t*>(hb + l.bh), reinterpret_cast<int32_t*>(hb + l.kh), l.ksh);
    if (vert) axis_coefficients(in_h, out_h, reinterpret_cast<int32_t*>(hb + l.bv), reinterpret_cast<int32_t*>(hb + l.kv), l.ksv);
    // pageable source: the call returns once the table is staged, so `host` may go out of scope
    UMV_CUDA_OK(cudaMemcpyAsync(ws, hb, l.tmp, cudaMemcpyHostToDevice, st));
    const uint8_t* mid = src;
    if (horiz) {
        uint8_t* o = vert ? ws + l.tmp : dst;
        launch_k(resample_kernel<true>, dim3((out_w + 127) / 128, in_h), dim3(128), 0, st, src, in_w, in_h, out_w,
                 reinterpret_cast<const int32_t*>(ws + l.bh), reinterpret_cast<const int32_t*>(ws + l.kh), l.ksh, o);
        UMV_LAUNCH_CHECK("resample_kernel<horizontal>");
        mid = o;
    }
    if (vert) {
        launch_k(resample_kernel<false>, dim3((out_w + 127) / 128, out_h), dim3(128), 0, st, mid, out_w, out_h, out_w,
                 reinterpret_cast<const int32_t*>(ws + l.bv), reinterpret_cast<const int32_t*>(ws + l.kv), l.ksv, dst);
        UMV_LAUNCH_CHECK("resample_kernel<vertical>");
    }
    return UMV_OK;
}

}  // extern "C"
