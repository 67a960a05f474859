// FLUX-style KL autoencoder (AutoEncoder, autoencoder.py:38-307) on NHWC bf16 activations.
// 3x3 convolutions = tap-major im2col (nearest-2x upsample and the stride-2 asymmetric pad folded into the
// gather) + the tcgen05 token-major linear kernel (bias / residual epilogues); GroupNorm(32)+swish computed
// in fp32 as under CUDA autocast and rounded once at the conv input; the single-head d=512 mid attention
// as fp32 scores -> fp32 softmax -> bf16 P -> tcgen05 PV.  Replaces the cuDNN / native_group_norm / SDPA
// call sites of SURVEY.md section 2.2 K17-K18.
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "engine.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
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
#define UMV_TRY(expr)            \
    do {                         \
        int _rc = (expr);        \
        if (_rc != UMV_OK) return _rc; \
    } while (0)

// ------------------------------------------------------------------------------------- layout
// NCHW bf16 -> NHWC bf16 with an optional affine (AutoEncoder.decode: z / scale + shift, two bf16 roundings).
__global__ void nchw_to_nhwc_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int C, int HW, float inv_scale,
                                    float shift, int affine) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * HW) return;
    const int p = i / C, c = i % C;
    float v = b2f(x[(size_t)c * HW + p]);
    if (affine) v = rbf(rbf(v / inv_scale) + shift);     // inv_scale carries scale_factor: z / 0.3611 + 0.1159
    y[i] = f2b(v);
}
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int C, int HW) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * HW) return;
    const int c = i / HW, p = i % HW;
    y[i] = x[(size_t)p * C + c];
}

// Packed latent tokens (fp32 [h*w, p*p*C], token = latent patch, channel-minor "hw pqc") -> the NHWC latent image the decoder reads,
// i.e. InterleaveInferencer.decode_image's reshape + einsum("nhwpqc->nchpwq") + .to(bf16) (inferencer.py:239-249) followed by
// AutoEncoder.decode's z / scale + shift (two bf16 roundings).
__global__ void tokens_to_nhwc_kernel(const float* __restrict__ tok, bf16* __restrict__ y, int h, int w, int p, int C, float scale,
                                      float shift) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int Wl = w * p;
    if (i >= h * p * Wl * C) return;
    const int c = i % C, px = (i / C) % Wl, py = i / (C * Wl);
    const int hh = py / p, pp = py % p, ww = px / p, q = px % p;
    float v = rbf(tok[((size_t)(hh * w + ww) * p * p + (pp * p + q)) * C + c]);
    y[i] = f2b(rbf(rbf(v / scale) + shift));
}
// (image * 0.5 + 0.5).clamp(0, 1) * 255 -> uint8 (inferencer.py:253-254) on the decoder's NHWC bf16 output: every torch op on the bf16
// tensor rounds to bf16 (x * 0.5 is exact), the cast to uint8 truncates.  HWC uint8 out -- the layout Image.fromarray takes.
__global__ void image_to_u8_kernel(const bf16* __restrict__ x, uint8_t* __restrict__ y, int n) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = rbf(b2f(x[i]) * 0.5f);
    v = rbf(v + 0.5f);
    v = fminf(fmaxf(v, 0.f), 1.f);
    v = rbf(v * 255.f);
    y[i] = (uint8_t)v;
}
// DiagonalGaussian.forward + AutoEncoder.encode's affine (autoencoder.py:266-272,300-303) on bf16 tensors, op by op:
// std = exp(0.5 * logvar); z = mean + std * noise; out = scale * (z - shift).  moments [2C, HW] (mean | logvar), noise / out [C, HW].
__global__ void vae_sample_kernel(const bf16* __restrict__ mom, const bf16* __restrict__ noise, bf16* __restrict__ out, int CHW,
                                  float scale, float shift) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CHW) return;
    float z = b2f(mom[i]);
    if (noise) {
        const float sd = rbf(expf(rbf(0.5f * b2f(mom[CHW + i]))));
        z = rbf(z + rbf(sd * b2f(noise[i])));
    }
    out[i] = f2b(rbf(scale * rbf(z - shift)));
}
// Patchify the encoded latent for the LLM (bagel.py:760-765: latent[:, :h*p, :w*p].reshape(C, h, p, w, p) -> "chpwq->hwpqc"):
// z bf16 [C, Hl, Wl] (padded batch plane) -> rows [h*w, p*p*C].
__global__ void latent_patchify_kernel(const bf16* __restrict__ z, bf16* __restrict__ rows, int C, int Hl, int Wl, int h, int w, int p) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = p * p * C;
    if (i >= h * w * per) return;
    const int c = i % C, q = (i / C) % p, pp = (i / (C * p)) % p, t = i / per;
    const int hh = t / w, ww = t % w;
    rows[i] = z[((size_t)c * Hl + (hh * p + pp)) * Wl + (ww * p + q)];
}

// im2col for a 3x3 conv on NHWC: out[(y*Wo+x), (ky*3+kx)*C + c].  up: source is the nearest-2x upsampled input
// (Upsample, autoencoder.py:116-119); stride 2: Downsample's pad (0,1,0,1) + stride-2 valid conv (:104-108).
__global__ void im2col3x3_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int H, int W, int C, int Ho, int Wo,
                                 int up, int stride) {
    pdl_launch_dependents();
    pdl_wait();
    const int pos = blockIdx.x;                  // output pixel
    const int oy = pos / Wo, ox = pos % Wo;
    const int chunks = C / 8;
    for (int i = threadIdx.x; i < 9 * chunks; i += blockDim.x) {
        const int tap = i / chunks, ch = i % chunks;
        const int ky = tap / 3, kx = tap % 3;
        int iy, ix;
        if (stride == 2) { iy = oy * 2 + ky; ix = ox * 2 + kx; }
        else { iy = oy + ky - 1; ix = ox + kx - 1; }
        const int Hs = up ? H * 2 : H, Ws = up ? W * 2 : W;    // extent of the (virtual) conv input
        U4 v = {0, 0, 0, 0};
        if (iy >= 0 && iy < Hs && ix >= 0 && ix < Ws) {
            const int sy = up ? iy >> 1 : iy, sx = up ? ix >> 1 : ix;
            v = ldg16(x + ((size_t)sy * W + sx) * C + ch * 8);
        }
        stg16(col + (size_t)pos * 9 * C + (size_t)tap * C + ch * 8, v);
    }
}
// generic (C not a multiple of 8, e.g. the 3-channel encoder input)
__global__ void im2col3x3_small_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int H, int W, int C, int Ho, int Wo) {
    pdl_launch_dependents();
    pdl_wait();
    const int pos = blockIdx.x;
    const int oy = pos / Wo, ox = pos % Wo;
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
        const int tap = i / C, c = i % C;
        const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
        bf16 v = f2b(0.f);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[((size_t)iy * W + ix) * C + c];
        col[(size_t)pos * 9 * C + i] = v;
    }
}

// ------------------------------------------------------------------------------------- GroupNorm
// partial sums per (row chunk, group): deterministic two-stage reduction (no atomics)
constexpr int kGnRows = 64;
__global__ void __launch_bounds__(256) gn_partial_kernel(const bf16* __restrict__ x, float* __restrict__ part, int HW, int C,
                                                          int groups) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float ssum[32 * 8], ssq[32 * 8];
    const int chunk = blockIdx.x;
    const int r0 = chunk * kGnRows, r1 = min(HW, r0 + kGnRows);
    const int cg = C / groups;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // thread t owns channels t, t+256, ... ; a group is cg consecutive channels
    for (int g = threadIdx.x; g < groups * 8; g += blockDim.x) { ssum[g] = 0.f; ssq[g] = 0.f; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f, q = 0.f;
        for (int r = r0; r < r1; ++r) {
            const float v = b2f(x[(size_t)r * C + c]);
            s += v;
            q += v * v;
        }
        // combine the cg channels of a group in a fixed order: slot (c % cg) < 8 when cg <= 8, else serial adds
        const int g = c / cg, j = c % cg;
        if (cg <= 8) { ssum[g * 8 + j] = s; ssq[g * 8 + j] = q; }
        else { atomicAdd(&ssum[g * 8 + (j & 7)], s); atomicAdd(&ssq[g * 8 + (j & 7)], q); }
    }
    __syncthreads();
    (void)warp; (void)lane;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        float s = 0.f, q = 0.f;
        for (int j = 0; j < 8; ++j) { s += ssum[g * 8 + j]; q += ssq[g * 8 + j]; }
        part[((size_t)chunk * groups + g) * 2] = s;
        part[((size_t)chunk * groups + g) * 2 + 1] = q;
    }
}
__global__ void gn_finalize_kernel(const float* __restrict__ part, float* __restrict__ stats, int chunks, int groups, int count,
                                   float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    double s = 0.0, q = 0.0;
    for (int c = 0; c < chunks; ++c) {
        s += part[((size_t)c * groups + g) * 2];
        q += part[((size_t)c * groups + g) * 2 + 1];
    }
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0) var = 0;
    stats[g * 2] = (float)mean;
    stats[g * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}
// y = bf16(swish((x - mean) * rstd * w + b)) in fp32 (CUDA autocast: GroupNorm and the swish after it are fp32,
// the conv that follows rounds its input once)
__global__ void gn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ stats, const bf16* __restrict__ w,
                                const bf16* __restrict__ b, bf16* __restrict__ y, int C, int groups, int act, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % C);
    const int g = c / (C / groups);
    float v = (b2f(x[i]) - stats[g * 2]) * stats[g * 2 + 1] * b2f(w[c]) + b2f(b[c]);
    if (act) v = v * (1.0f / (1.0f + expf(-v)));
    y[i] = f2b(v);
}

// ------------------------------------------------------------------------------------- mid attention
// S[i][j] = scale * q_i . k_j  (fp32), 32x32 tiles
__global__ void __launch_bounds__(256) attn_scores_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k,
                                                           float* __restrict__ s, int L, int C, float scale) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sq[32][33], sk[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < C; c0 += 32) {
        for (int r = ty; r < 32; r += 8) {
            sq[r][tx] = (i0 + r < L && c0 + tx < C) ? b2f(q[(size_t)(i0 + r) * C + c0 + tx]) : 0.f;
            sk[r][tx] = (j0 + r < L && c0 + tx < C) ? b2f(k[(size_t)(j0 + r) * C + c0 + tx]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = ty + a * 8;
#pragma unroll 8
            for (int c = 0; c < 32; ++c) acc[a] = fmaf(sq[i][c], sk[tx][c], acc[a]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty + a * 8, j = j0 + tx;
        if (i < L && j < L) s[(size_t)i * L + j] = acc[a] * scale;
    }
}
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, bf16* __restrict__ p, int L) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[32];
    const float* row = s + (size_t)blockIdx.x * L;
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < L; j += blockDim.x) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
    float sum = 0.f;
    for (int j = threadIdx.x; j < L; j += blockDim.x) sum += expf(row[j] - mx);
    sum = block_sum(sum, red);
    const float inv = 1.0f / sum;
    for (int j = threadIdx.x; j < L; j += blockDim.x) p[(size_t)blockIdx.x * L + j] = f2b(expf(row[j] - mx) * inv);
}
__global__ void transpose_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int R, int Cc) {   // [R,Cc] -> [Cc,R]
    pdl_launch_dependents();
    pdl_wait();
    __shared__ bf16 t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8)
        if (r0 + r < R && c0 + tx < Cc) t[r][tx] = x[(size_t)(r0 + r) * Cc + c0 + tx];
    __syncthreads();
    for (int c = ty; c < 32; c += 8)
        if (c0 + c < Cc && r0 + tx < R) y[(size_t)(c0 + c) * R + r0 + tx] = t[tx][c];
}

// ------------------------------------------------------------------------------------- build
static int reg_conv(umv_engine* e, const std::string& name, VaeConv* c, int cout, int cin, int k) {
    c->cin = cin; c->cout = cout; c->k = k;
    void* p = nullptr;
    UMV_TRY(engine_alloc(e, &p, (size_t)cout * cin * k * k * 2));
    c->w = static_cast<bf16*>(p);
    UMV_TRY(engine_alloc(e, &p, (size_t)cout * 2));
    c->b = static_cast<bf16*>(p);
    const float bound = 1.0f / sqrtf((float)(cin * k * k));
    engine_reg(e, name + ".weight", c->w, cout, (int64_t)cin * k * k, 4, k, cin, bound, 0.f);
    engine_reg(e, name + ".bias", c->b, 1, cout, 1, 0, 0, bound, 0.f);
    return UMV_OK;
}
static int reg_norm(umv_engine* e, const std::string& name, VaeNorm* n, int c) {
    n->c = c;
    void* p = nullptr;
    UMV_TRY(engine_alloc(e, &p, (size_t)c * 2));
    n->w = static_cast<bf16*>(p);
    UMV_TRY(engine_alloc(e, &p, (size_t)c * 2));
    n->b = static_cast<bf16*>(p);
    engine_reg(e, name + ".weight", n->w, 1, c, 1, 0, 0, 0.1f, 1.f);
    engine_reg(e, name + ".bias", n->b, 1, c, 1, 0, 0, 0.05f, 0.f);
    return UMV_OK;
}
static int reg_res(umv_engine* e, const std::string& p, VaeRes* r, int cin, int cout) {
    r->cin = cin; r->cout = cout;
    UMV_TRY(reg_norm(e, p + "norm1", &r->n1, cin));
    UMV_TRY(reg_conv(e, p + "conv1", &r->c1, cout, cin, 3));
    UMV_TRY(reg_norm(e, p + "norm2", &r->n2, cout));
    UMV_TRY(reg_conv(e, p + "conv2", &r->c2, cout, cout, 3));
    if (cin != cout) UMV_TRY(reg_conv(e, p + "nin_shortcut", &r->sc, cout, cin, 1));
    return UMV_OK;
}
static int reg_attn(umv_engine* e, const std::string& p, VaeAttn* a, int c) {
    a->c = c;
    UMV_TRY(reg_norm(e, p + "norm", &a->n, c));
    UMV_TRY(reg_conv(e, p + "q", &a->q, c, c, 1));
    UMV_TRY(reg_conv(e, p + "k", &a->k, c, c, 1));
    UMV_TRY(reg_conv(e, p + "v", &a->v, c, c, 1));
    UMV_TRY(reg_conv(e, p + "proj_out", &a->o, c, c, 1));
    return UMV_OK;
}

int vae_build(umv_engine* e) {
    e->vae = new VaeState();
    VaeState& V = *e->vae;
    const int ch = V.ch, z = V.z, nlev = V.nlev;
    {   // Encoder (autoencoder.py:122-167)
        const std::string P = "vae_model.encoder.";
        VaeHalf& H = V.enc;
        UMV_TRY(reg_conv(e, P + "conv_in", &H.conv_in, ch, 3, 3));
        H.levels.resize(nlev);
        int block_in = ch;
        for (int l = 0; l < nlev; ++l) {
            block_in = ch * (l == 0 ? 1 : V.mult[l - 1]);
            const int block_out = ch * V.mult[l];
            for (int b = 0; b < V.nres; ++b) {
                H.levels[l].blocks.emplace_back();
                UMV_TRY(reg_res(e, P + "down." + std::to_string(l) + ".block." + std::to_string(b) + ".", &H.levels[l].blocks.back(),
                                block_in, block_out));
                block_in = block_out;
            }
            if (l != nlev - 1) {
                H.levels[l].has_resample = true;
                UMV_TRY(reg_conv(e, P + "down." + std::to_string(l) + ".downsample.conv", &H.levels[l].resample, block_in, block_in, 3));
            }
        }
        UMV_TRY(reg_res(e, P + "mid.block_1.", &H.mid1, block_in, block_in));
        UMV_TRY(reg_attn(e, P + "mid.attn_1.", &H.attn, block_in));
        UMV_TRY(reg_res(e, P + "mid.block_2.", &H.mid2, block_in, block_in));
        UMV_TRY(reg_norm(e, P + "norm_out", &H.norm_out, block_in));
        UMV_TRY(reg_conv(e, P + "conv_out", &H.conv_out, 2 * z, block_in, 3));
    }
    {   // Decoder (autoencoder.py:190-238)
        const std::string P = "vae_model.decoder.";
        VaeHalf& H = V.dec;
        int block_in = ch * V.mult[nlev - 1];
        UMV_TRY(reg_conv(e, P + "conv_in", &H.conv_in, block_in, z, 3));
        UMV_TRY(reg_res(e, P + "mid.block_1.", &H.mid1, block_in, block_in));
        UMV_TRY(reg_attn(e, P + "mid.attn_1.", &H.attn, block_in));
        UMV_TRY(reg_res(e, P + "mid.block_2.", &H.mid2, block_in, block_in));
        H.levels.resize(nlev);
        for (int l = nlev - 1; l >= 0; --l) {
            const int block_out = ch * V.mult[l];
            for (int b = 0; b < V.nres + 1; ++b) {
                H.levels[l].blocks.emplace_back();
                UMV_TRY(reg_res(e, P + "up." + std::to_string(l) + ".block." + std::to_string(b) + ".", &H.levels[l].blocks.back(),
                                block_in, block_out));
                block_in = block_out;
            }
            if (l != 0) {
                H.levels[l].has_resample = true;
                UMV_TRY(reg_conv(e, P + "up." + std::to_string(l) + ".upsample.conv", &H.levels[l].resample, block_in, block_in, 3));
            }
        }
        UMV_TRY(reg_norm(e, P + "norm_out", &H.norm_out, block_in));
        UMV_TRY(reg_conv(e, P + "conv_out", &H.conv_out, 3, block_in, 3));
    }
    return UMV_OK;
}

// ------------------------------------------------------------------------------------- forward
static int vae_reserve(umv_engine* e, size_t act, size_t col, size_t s_elems) {
    VaeState& V = *e->vae;
    void* p;
    if (act > V.act_elems) {
        UMV_TRY(engine_alloc(e, &p, act * 2)); V.a0 = static_cast<bf16*>(p);
        UMV_TRY(engine_alloc(e, &p, act * 2)); V.a1 = static_cast<bf16*>(p);
        UMV_TRY(engine_alloc(e, &p, act * 2)); V.a2 = static_cast<bf16*>(p);
        V.act_elems = act;
    }
    if (col > V.col_elems) { UMV_TRY(engine_alloc(e, &p, col * 2)); V.col = static_cast<bf16*>(p); V.col_elems = col; }
    if (s_elems > V.s_elems) {
        UMV_TRY(engine_alloc(e, &p, s_elems * 4)); V.s = static_cast<float*>(p);
        UMV_TRY(engine_alloc(e, &p, s_elems * 2)); V.p = static_cast<bf16*>(p);
        V.s_elems = s_elems;
    }
    if (!V.stats) {
        UMV_TRY(engine_alloc(e, &p, (size_t)(1 << 22)));     // GroupNorm partials: chunks * 32 groups * 2 floats
        V.stats = static_cast<float*>(p);
        UMV_TRY(engine_alloc(e, &p, (size_t)4096 * 512 * 2)); // V^T for the mid attention (L <= 4096, C = 512)
        V.vt = static_cast<bf16*>(p);
    }
    return UMV_OK;
}

struct Img { int H, W; };

static int conv(umv_engine* e, const VaeConv& c, const bf16* x, Img in, bf16* y, Img* out, int up, int stride,
                const bf16* residual, cudaStream_t st) {
    VaeState& V = *e->vae;
    Img o = in;
    if (up) { o.H *= 2; o.W *= 2; }
    if (stride == 2) { o.H = in.H / 2; o.W = in.W / 2; }
    const int M = o.H * o.W;
    const int epi = residual ? EPI_RESID : EPI_BF16;
    if (c.k == 1) {
        UMV_TRY(lin(e, x, c.cin, c.w, c.b, residual, y, c.cout, M, c.cout, c.cin, epi, st));
    } else {
        if (c.cin % 8 == 0) {
            launch_k(im2col3x3_kernel, dim3(M), dim3(128), 0, st, x, V.col, in.H, in.W, c.cin, o.H, o.W, up, stride);
            UMV_LAUNCH_CHECK("im2col3x3_kernel");
        } else {
            launch_k(im2col3x3_small_kernel, dim3(M), dim3(32), 0, st, x, V.col, in.H, in.W, c.cin, o.H, o.W);
            UMV_LAUNCH_CHECK("im2col3x3_small_kernel");
        }
        UMV_TRY(lin(e, V.col, 9 * c.cin, c.w, c.b, residual, y, c.cout, M, c.cout, 9 * c.cin, epi, st));
    }
    if (out) *out = o;
    return UMV_OK;
}

static int gn(umv_engine* e, const VaeNorm& n, const bf16* x, Img im, bf16* y, int act, cudaStream_t st) {
    VaeState& V = *e->vae;
    const int HW = im.H * im.W, C = n.c, groups = 32;
    const int chunks = (HW + kGnRows - 1) / kGnRows;
    float* part = V.stats + 64;       // stats[0..63] = (mean, rstd) per group
    UMV_REQUIRE((size_t)chunks * groups * 2 * 4 + 256 <= (size_t)(1 << 22), UMV_ERR_NOMEM, "GroupNorm: image too large (%d px)", HW);
    launch_k(gn_partial_kernel, dim3(chunks), dim3(256), 0, st, x, part, HW, C, groups);
    UMV_LAUNCH_CHECK("gn_partial_kernel");
    launch_k(gn_finalize_kernel, dim3(1), dim3(32), 0, st, (const float*)part, V.stats, chunks, groups, HW * (C / groups), 1e-6f);
    UMV_LAUNCH_CHECK("gn_finalize_kernel");
    const size_t nel = (size_t)HW * C;
    launch_k(gn_apply_kernel, dim3((unsigned)((nel + 255) / 256)), dim3(256), 0, st, x, (const float*)V.stats, n.w, n.b, y, C, groups, act,
             nel);
    UMV_LAUNCH_CHECK("gn_apply_kernel");
    return UMV_OK;
}

// ResnetBlock (autoencoder.py:82-95): x + conv2(swish(norm2(conv1(swish(norm1(x))))))  (nin_shortcut on x if cin != cout)
// in: V.a0 -> out: V.a0 (uses a1, a2)
static int res_block(umv_engine* e, const VaeRes& r, Img im, cudaStream_t st) {
    VaeState& V = *e->vae;
    UMV_TRY(gn(e, r.n1, V.a0, im, V.a1, 1, st));
    UMV_TRY(conv(e, r.c1, V.a1, im, V.a2, nullptr, 0, 1, nullptr, st));
    UMV_TRY(gn(e, r.n2, V.a2, im, V.a1, 1, st));
    const bf16* skip = V.a0;
    if (r.cin != r.cout) {
        UMV_TRY(conv(e, r.sc, V.a0, im, V.a2, nullptr, 0, 1, nullptr, st));
        skip = V.a2;
    }
    // conv2 with the residual folded into the epilogue: bf16(bf16(acc + bias) + skip); output may alias a0 only
    // when skip != a0 rows are read by the same thread that writes them (true for the epilogue), so write to a0.
    bf16* dst = V.a0;
    if (skip == V.a0) {
        UMV_TRY(conv(e, r.c2, V.a1, im, dst, nullptr, 0, 1, skip, st));
    } else {
        UMV_TRY(conv(e, r.c2, V.a1, im, dst, nullptr, 0, 1, skip, st));
    }
    return UMV_OK;
}

// AttnBlock (autoencoder.py:50-65), in/out: V.a0
static int attn_block(umv_engine* e, const VaeAttn& a, Img im, cudaStream_t st) {
    VaeState& V = *e->vae;
    const int L = im.H * im.W, C = a.c;
    UMV_REQUIRE(L <= 4096 && C <= 512, UMV_ERR_UNSUPPORTED, "VAE attention: %d positions x %d channels exceeds the workspace", L, C);
    UMV_TRY(gn(e, a.n, V.a0, im, V.a1, 0, st));
    bf16* q = V.a2;
    bf16* k = V.a2 + (size_t)L * C;
    bf16* v = V.a2 + (size_t)2 * L * C;
    UMV_TRY(lin(e, V.a1, C, a.q.w, a.q.b, nullptr, q, C, L, C, C, EPI_BF16, st));
    UMV_TRY(lin(e, V.a1, C, a.k.w, a.k.b, nullptr, k, C, L, C, C, EPI_BF16, st));
    UMV_TRY(lin(e, V.a1, C, a.v.w, a.v.b, nullptr, v, C, L, C, C, EPI_BF16, st));
    launch_k(attn_scores_kernel, dim3((L + 31) / 32, (L + 31) / 32), dim3(256), 0, st, (const bf16*)q, (const bf16*)k, V.s, L, C,
             1.0f / sqrtf((float)C));
    UMV_LAUNCH_CHECK("attn_scores_kernel");
    launch_k(softmax_rows_kernel, dim3(L), dim3(256), 0, st, (const float*)V.s, V.p, L);
    UMV_LAUNCH_CHECK("softmax_rows_kernel");
    launch_k(transpose_kernel, dim3((C + 31) / 32, (L + 31) / 32), dim3(256), 0, st, (const bf16*)v, V.vt, L, C);
    UMV_LAUNCH_CHECK("transpose_kernel");
    {   // O = P V: the "weight" operand V^T was written by the transpose just above -> no early weight prefetch
        LinearCall c;
        c.x = V.p; c.ldx = L; c.w = V.vt; c.y = V.a1; c.ldy = C; c.M = L; c.N = C; c.K = L; c.epi = EPI_BF16; c.w_static = false;
        UMV_TRY(linear_forward(c, st));
    }
    return lin(e, V.a1, C, a.o.w, a.o.b, V.a0, V.a0, C, L, C, C, EPI_RESID, st);               // x + proj_out(O)
}

static size_t max_act_elems(int H8, int W8, bool decode) {
    // largest activation: full resolution x 256 channels right after the last upsample conv (decoder) or
    // full resolution x 128 (encoder); 3 q/k/v planes at the bottleneck
    const size_t full = (size_t)H8 * 8 * W8 * 8;
    return std::max(full * (decode ? 256 : 128), (size_t)3 * H8 * W8 * 512) + 64;
}

// Decoder.forward (autoencoder.py:240-257) on the NHWC latent in V.a1 (already z / scale + shift); leaves conv_out's NHWC
// [H*W, 3] bf16 image in V.a2 and the output geometry in *im.
static int decode_body(umv_engine* e, Img* im_io, cudaStream_t st) {
    VaeState& V = *e->vae;
    const VaeHalf& H = V.dec;
    Img im = *im_io;
    UMV_TRY(conv(e, H.conv_in, V.a1, im, V.a0, nullptr, 0, 1, nullptr, st));
    UMV_TRY(res_block(e, H.mid1, im, st));
    UMV_TRY(attn_block(e, H.attn, im, st));
    UMV_TRY(res_block(e, H.mid2, im, st));
    for (int l = V.nlev - 1; l >= 0; --l) {
        for (const VaeRes& r : H.levels[l].blocks) UMV_TRY(res_block(e, r, im, st));
        if (H.levels[l].has_resample) {
            UMV_TRY(conv(e, H.levels[l].resample, V.a0, im, V.a1, &im, 1, 1, nullptr, st));   // nearest x2 folded into the gather
            std::swap(V.a0, V.a1);
        }
    }
    UMV_TRY(gn(e, H.norm_out, V.a0, im, V.a1, 1, st));
    UMV_TRY(conv(e, H.conv_out, V.a1, im, V.a2, nullptr, 0, 1, nullptr, st));
    *im_io = im;
    return UMV_OK;
}

int latent_patchify(const bf16* z, bf16* rows, int C, int Hl, int Wl, int h, int w, int p, cudaStream_t st) {
    const int n = h * w * p * p * C;
    launch_k(latent_patchify_kernel, dim3((n + 255) / 256), dim3(256), 0, st, z, rows, C, Hl, Wl, h, w, p);
    UMV_LAUNCH_CHECK("latent_patchify_kernel");
    return UMV_OK;
}

}  // namespace umv

using namespace umv;

extern "C" {

int umv_vae_decode(umv_engine* e, const void* z, int32_t n, int32_t h, int32_t w, void* out, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->vae, UMV_ERR_STATE, "VAE weights were not enabled");
    UMV_REQUIRE(z && out && n > 0 && h > 0 && w > 0, UMV_ERR_INVALID, "umv_vae_decode: null/empty argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VaeState& V = *e->vae;
    const size_t full = (size_t)h * 8 * w * 8;
    UMV_TRY(vae_reserve(e, max_act_elems(h, w, true), full * 9 * 256 + 64, (size_t)h * w * h * w));
    for (int i = 0; i < n; ++i) {
        const bf16* zi = static_cast<const bf16*>(z) + (size_t)i * V.z * h * w;
        const int nel = V.z * h * w;
        launch_k(nchw_to_nhwc_kernel, dim3((nel + 255) / 256), dim3(256), 0, st, zi, V.a1, V.z, h * w, V.scale, V.shift, 1);
        UMV_LAUNCH_CHECK("nchw_to_nhwc_kernel");
        Img im{h, w};
        UMV_TRY(decode_body(e, &im, st));
        const int HW = im.H * im.W;
        launch_k(nhwc_to_nchw_kernel, dim3((3 * HW + 255) / 256), dim3(256), 0, st, (const bf16*)V.a2,
                 static_cast<bf16*>(out) + (size_t)i * 3 * HW, 3, HW);
        UMV_LAUNCH_CHECK("nhwc_to_nchw_kernel");
    }
    return UMV_OK;
}

// Parity hook: ONE block of the autoencoder on a caller-provided activation (block-level tests against the oracle / the reference's
// own sub-modules).  path: the reference's module path inside AutoEncoder, e.g. "decoder.mid.block_1", "decoder.mid.attn_1",
// "decoder.up.3.block.0", "decoder.up.2.upsample" (nearest x2 + conv), "encoder.down.0.downsample", "decoder.conv_in",
// "decoder.norm_out" (GroupNorm + swish).  x / out: bf16 [C, H, W] (one image, NCHW as the reference); out_chw receives the output shape.
int umv_op_vae_block(umv_engine* e, const char* path, const void* x, int32_t C, int32_t Hh, int32_t Ww, void* out, int32_t* out_chw,
                     void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->vae, UMV_ERR_STATE, "VAE weights were not enabled");
    UMV_REQUIRE(path && x && out && out_chw && C > 0 && Hh > 0 && Ww > 0, UMV_ERR_INVALID, "umv_op_vae_block: null/empty argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VaeState& V = *e->vae;
    const std::string p(path);
    const bool dec = p.rfind("decoder.", 0) == 0;
    UMV_REQUIRE(dec || p.rfind("encoder.", 0) == 0, UMV_ERR_INVALID, "umv_op_vae_block: path must start with encoder. / decoder.");
    const VaeHalf& Hf = dec ? V.dec : V.enc;
    const std::string rest = p.substr(8);
    const size_t px = (size_t)Hh * Ww;
    UMV_TRY(vae_reserve(e, px * 4 * 512 + 64, px * 4 * 9 * 512 + 64, px <= 4096 ? px * px : 1));
    launch_k(nchw_to_nhwc_kernel, dim3((unsigned)((C * px + 255) / 256)), dim3(256), 0, st, static_cast<const bf16*>(x), V.a0, C, (int)px, 1.f, 0.f, 0);
    UMV_LAUNCH_CHECK("nchw_to_nhwc_kernel");
    Img im{Hh, Ww};
    const bf16* res = V.a0;
    int Cout = C;
    auto res_at = [&](const VaeRes& r) -> int {
        UMV_REQUIRE(r.cin == C, UMV_ERR_INVALID, "umv_op_vae_block: '%s' takes %d channels, got %d", path, r.cin, C);
        Cout = r.cout;
        return res_block(e, r, im, st);
    };
    int lvl = -1, bi = -1;
    if (rest == "mid.block_1") { UMV_TRY(res_at(Hf.mid1)); }
    else if (rest == "mid.block_2") { UMV_TRY(res_at(Hf.mid2)); }
    else if (rest == "mid.attn_1") {
        UMV_REQUIRE(Hf.attn.c == C, UMV_ERR_INVALID, "umv_op_vae_block: attention takes %d channels", Hf.attn.c);
        UMV_TRY(attn_block(e, Hf.attn, im, st));
    } else if (sscanf(rest.c_str(), dec ? "up.%d.block.%d" : "down.%d.block.%d", &lvl, &bi) == 2) {
        UMV_REQUIRE(lvl >= 0 && lvl < V.nlev && bi >= 0 && bi < (int)Hf.levels[lvl].blocks.size(), UMV_ERR_INVALID, "no block '%s'", path);
        UMV_TRY(res_at(Hf.levels[lvl].blocks[bi]));
    } else if (sscanf(rest.c_str(), dec ? "up.%d.upsample" : "down.%d.downsample", &lvl) == 1) {
        UMV_REQUIRE(lvl >= 0 && lvl < V.nlev && Hf.levels[lvl].has_resample && Hf.levels[lvl].resample.cin == C, UMV_ERR_INVALID,
                    "no resample '%s' for %d channels", path, C);
        UMV_TRY(conv(e, Hf.levels[lvl].resample, V.a0, im, V.a1, &im, dec ? 1 : 0, dec ? 1 : 2, nullptr, st));
        Cout = Hf.levels[lvl].resample.cout;
        res = V.a1;
    } else if (rest == "conv_in" || rest == "conv_out") {
        const VaeConv& c = rest == "conv_in" ? Hf.conv_in : Hf.conv_out;
        UMV_REQUIRE(c.cin == C, UMV_ERR_INVALID, "umv_op_vae_block: '%s' takes %d channels", path, c.cin);
        UMV_TRY(conv(e, c, V.a0, im, V.a1, nullptr, 0, 1, nullptr, st));
        Cout = c.cout;
        res = V.a1;
    } else if (rest == "norm_out") {
        UMV_REQUIRE(Hf.norm_out.c == C, UMV_ERR_INVALID, "umv_op_vae_block: norm_out takes %d channels", Hf.norm_out.c);
        UMV_TRY(gn(e, Hf.norm_out, V.a0, im, V.a1, 1, st));
        res = V.a1;
    } else {
        set_error("umv_op_vae_block: unknown block '%s'", path);
        return UMV_ERR_INVALID;
    }
    const int HW = im.H * im.W;
    launch_k(nhwc_to_nchw_kernel, dim3((Cout * HW + 255) / 256), dim3(256), 0, st, res, static_cast<bf16*>(out), Cout, HW);
    UMV_LAUNCH_CHECK("nhwc_to_nchw_kernel");
    out_chw[0] = Cout; out_chw[1] = im.H; out_chw[2] = im.W;
    return UMV_OK;
}

int umv_decode_image_u8(umv_engine* e, const float* latent_tokens, int32_t n, int32_t h, int32_t w, int32_t p, uint8_t* out,
                        void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->vae, UMV_ERR_STATE, "VAE weights were not enabled");
    UMV_REQUIRE(latent_tokens && out && n > 0 && h > 0 && w > 0 && p > 0, UMV_ERR_INVALID, "umv_decode_image_u8: null/empty argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VaeState& V = *e->vae;
    UMV_REQUIRE(p * p * V.z == e->d.latent_dim, UMV_ERR_INVALID, "umv_decode_image_u8: latent patch %d does not match latent_dim %d", p,
                e->d.latent_dim);
    const int hl = h * p, wl = w * p;
    const size_t full = (size_t)hl * 8 * wl * 8;
    UMV_TRY(vae_reserve(e, max_act_elems(hl, wl, true), full * 9 * 256 + 64, (size_t)hl * wl * hl * wl));
    for (int i = 0; i < n; ++i) {
        const int nel = V.z * hl * wl;
        launch_k(tokens_to_nhwc_kernel, dim3((nel + 255) / 256), dim3(256), 0, st, latent_tokens + (size_t)i * nel, V.a1, h, w, p, V.z,
                 V.scale, V.shift);
        UMV_LAUNCH_CHECK("tokens_to_nhwc_kernel");
        Img im{hl, wl};
        UMV_TRY(decode_body(e, &im, st));
        const int npx = 3 * im.H * im.W;
        launch_k(image_to_u8_kernel, dim3((npx + 255) / 256), dim3(256), 0, st, (const bf16*)V.a2, out + (size_t)i * npx, npx);
        UMV_LAUNCH_CHECK("image_to_u8_kernel");
    }
    return UMV_OK;
}

int umv_vae_sample(umv_engine* e, const void* moments, const void* noise, int32_t n, int32_t h, int32_t w, void* out, void* stream) {
    UMV_REQUIRE(e && e->vae, UMV_ERR_STATE, "VAE weights were not enabled");
    UMV_REQUIRE(moments && out && n > 0 && h > 0 && w > 0, UMV_ERR_INVALID, "umv_vae_sample: null/empty argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VaeState& V = *e->vae;
    const int chw = V.z * h * w;
    for (int i = 0; i < n; ++i) {
        launch_k(vae_sample_kernel, dim3((chw + 255) / 256), dim3(256), 0, st, static_cast<const bf16*>(moments) + (size_t)i * 2 * chw,
                 noise ? static_cast<const bf16*>(noise) + (size_t)i * chw : nullptr, static_cast<bf16*>(out) + (size_t)i * chw, chw,
                 V.scale, V.shift);
        UMV_LAUNCH_CHECK("vae_sample_kernel");
    }
    return UMV_OK;
}

int umv_vae_encode_moments(umv_engine* e, const void* x, int32_t n, int32_t Hh, int32_t Ww, void* out, void* stream) {
    UMV_REQUIRE(e && e->finalized, UMV_ERR_STATE, "engine not finalized");
    UMV_REQUIRE(e->vae, UMV_ERR_STATE, "VAE weights were not enabled");
    UMV_REQUIRE(x && out && n > 0 && Hh > 0 && Ww > 0 && Hh % 8 == 0 && Ww % 8 == 0, UMV_ERR_INVALID,
                "umv_vae_encode_moments: image sides must be positive multiples of 8");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VaeState& V = *e->vae;
    const int h = Hh / 8, w = Ww / 8;
    const size_t full = (size_t)Hh * Ww;
    UMV_TRY(vae_reserve(e, max_act_elems(h, w, false), full * 9 * 128 + 64, (size_t)h * w * h * w));
    const VaeHalf& H = V.enc;
    for (int i = 0; i < n; ++i) {
        const bf16* xi = static_cast<const bf16*>(x) + (size_t)i * 3 * Hh * Ww;
        Img im{Hh, Ww};
        launch_k(nchw_to_nhwc_kernel, dim3((unsigned)((3 * full + 255) / 256)), dim3(256), 0, st, xi, V.a1, 3, (int)full, 1.f, 0.f, 0);
        UMV_LAUNCH_CHECK("nchw_to_nhwc_kernel");
        UMV_TRY(conv(e, H.conv_in, V.a1, im, V.a0, nullptr, 0, 1, nullptr, st));
        for (int l = 0; l < V.nlev; ++l) {
            for (const VaeRes& r : H.levels[l].blocks) {
                UMV_TRY(res_block(e, r, im, st));
            }
            if (H.levels[l].has_resample) {
                UMV_TRY(conv(e, H.levels[l].resample, V.a0, im, V.a1, &im, 0, 2, nullptr, st));
                std::swap(V.a0, V.a1);
            }
        }
        UMV_TRY(res_block(e, H.mid1, im, st));
        UMV_TRY(attn_block(e, H.attn, im, st));
        UMV_TRY(res_block(e, H.mid2, im, st));
        UMV_TRY(gn(e, H.norm_out, V.a0, im, V.a1, 1, st));
        UMV_TRY(conv(e, H.conv_out, V.a1, im, V.a2, nullptr, 0, 1, nullptr, st));
        const int HW = im.H * im.W, C = 2 * V.z;
        launch_k(nhwc_to_nchw_kernel, dim3((C * HW + 255) / 256), dim3(256), 0, st, (const bf16*)V.a2,
                 static_cast<bf16*>(out) + (size_t)i * C * HW, C, HW);
        UMV_LAUNCH_CHECK("nhwc_to_nchw_kernel");
    }
    return UMV_OK;
}

}  // extern "C"
