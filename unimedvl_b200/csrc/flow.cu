// Rectified-flow glue kernels (Bagel._forward_flow / generate_image, bagel.py:901-1211;
// TimestepEmbedder modeling_utils.py:86-109): sinusoidal timestep features, latent-token composition
// (vae2llm(x_t) + t_emb + latent_pos_embed, markers from the embedding table), classifier-free-guidance
// mix + renorm with the reference's op-by-op bf16 roundings, and the Euler update.
// They replace ~15 eager elementwise kernels + ~20 gather/scatter per step.
#include "../../include/umv.h"
#include "common.cuh"
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

// t_freq = [cos(t f_i) | sin(t f_i)], fp32 math, handed to the autocast Linear as bf16.
__global__ void timestep_freq_kernel(float t, const float* __restrict__ freqs, int half, bf16* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < half) {
        const float a = __fmul_rn(t, freqs[i]);
        out[i] = f2b(cosf(a));
        out[half + i] = f2b(sinf(a));
    }
}
int timestep_freq(float t, const float* freqs, int half, bf16* out, cudaStream_t s) {
    launch_k(timestep_freq_kernel, dim3((half + 127) / 128), dim3(128), 0, s, t, freqs, half, out);
    UMV_LAUNCH_CHECK("timestep_freq_kernel");
    return UMV_OK;
}

__global__ void silu_rows_kernel(bf16* __restrict__ x, int n) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = f2b(silu_f(b2f(x[i])));
}
int silu_inplace(bf16* x, int n, cudaStream_t s) {
    launch_k(silu_rows_kernel, dim3((n + 255) / 256), dim3(256), 0, s, x, n);
    UMV_LAUNCH_CHECK("silu_rows_kernel");
    return UMV_OK;
}

// Packed query sequence of the flow step, replicated for every CFG branch:
// row_src[r] >= 0: latent token -> bf16(bf16(lat + t_emb) + pos_table[pos_id]);  -1 / -2: start / end marker embedding.
__global__ void flow_compose_kernel(const bf16* __restrict__ lat, const bf16* __restrict__ temb, const bf16* __restrict__ pos_table,
                                    const int64_t* __restrict__ pos_ids, const bf16* __restrict__ embed, int64_t id_start,
                                    int64_t id_end, const int* __restrict__ row_src, const int* __restrict__ row_dst,
                                    int rows_per_branch, int D, bf16* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x;
    const int src = row_src[r % rows_per_branch];
    bf16* dst = out + (size_t)(row_dst ? row_dst[r] : r) * D;      // row_dst: the forward's segregated row order (llm_run)
    if (src < 0) {
        const bf16* e = embed + (size_t)(src == -1 ? id_start : id_end) * D;
        for (int c = threadIdx.x; c < D / 8; c += blockDim.x) stg16(dst + c * 8, ldg16(e + c * 8));
        return;
    }
    const bf16* l = lat + (size_t)src * D;
    const bf16* p = pos_table + (size_t)pos_ids[src] * D;
    for (int c = threadIdx.x; c < D / 8; c += blockDim.x) {
        const U4 a = ldg16(l + c * 8), t = ldg16(temb + c * 8), q = ldg16(p + c * 8);
        const uint32_t *aw = &a.x, *tw = &t.x, *qw = &q.x;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = unpack2(aw[j]), ft = unpack2(tw[j]), fq = unpack2(qw[j]);
            o[j] = pack2(rbf(fa.x + ft.x) + fq.x, rbf(fa.y + ft.y) + fq.y);
        }
        stg16(dst + c * 8, U4{o[0], o[1], o[2], o[3]});
    }
}
int flow_compose(const bf16* lat, const bf16* temb, const bf16* pos_table, const int64_t* pos_ids, const bf16* embed,
                 int64_t id_start, int64_t id_end, const int* row_src, int rows_per_branch, int branches, int D, bf16* out,
                 cudaStream_t s, const int* row_dst) {
    launch_k(flow_compose_kernel, dim3(rows_per_branch * branches), dim3(128), 0, s, lat, temb, pos_table, pos_ids, embed,
             id_start, id_end, row_src, row_dst, rows_per_branch, D, out);
    UMV_LAUNCH_CHECK("flow_compose_kernel");
    return UMV_OK;
}

// CFG mix + renorm, one CTA per image, one warp per latent token (C <= 64 channels: lanes own c and c+32).
// v: bf16 [branches * rows_per_branch, C] (llm2vae output of every packed row); branch 0 = main context.
__device__ __forceinline__ float cfg_mix(float base, float scale, float v) {
    // base + scale * (v - base) on bf16 tensors: every op rounds (bagel.py:1175,1185,1189,1192)
    return rbf(base + rbf(scale * rbf(v - base)));
}

__global__ void __launch_bounds__(256) cfg_combine_kernel(CfgArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[32];
    const int img = blockIdx.x;
    const int n = a.img_n[img], row0 = a.img_row0[img], lat0 = a.img_lat0[img];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const bool has_text = a.text_branch >= 0, has_img = a.img_branch >= 0;
    const bf16* vt_p = a.v;
    const bf16* vx_p = a.v + (size_t)(has_text ? a.text_branch : 0) * a.rows_per_branch * a.C;
    const bf16* vi_p = a.v + (size_t)(has_img ? a.img_branch : 0) * a.rows_per_branch * a.C;

    float s0 = 0.f, s1 = 0.f;        // per-thread partial sums for the per-image ("global") norm
    for (int pass = 0; pass < 2; ++pass) {
        float gscale = 1.f;
        if (pass == 1) {
            if (!(has_text && a.renorm_type == 0)) break;
            const float t0 = block_sum(s0, red);
            const float t1 = block_sum(s1, red);
            gscale = fminf(fmaxf(sqrtf(t0) / (sqrtf(t1) + 1e-8f), a.renorm_min), 1.0f);
        }
        for (int tok = warp; tok < n; tok += nwarps) {
            const size_t base = (size_t)(row0 + tok) * a.C;
            float o[2];
            float vt[2], v_[2];
            float n0 = 0.f, n1 = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = lane + 32 * h;
                vt[h] = v_[h] = 0.f;
                if (c < a.C) {
                    vt[h] = b2f(vt_p[base + c]);
                    if (has_text) {
                        const float vx = b2f(vx_p[base + c]);
                        v_[h] = cfg_mix(vx, a.text_scale, vt[h]);                               // v_t_text_
                        if (a.renorm_type != 2 && has_img) v_[h] = cfg_mix(b2f(vi_p[base + c]), a.img_scale, v_[h]);
                    } else {
                        v_[h] = vt[h];
                    }
                }
                n0 += vt[h] * vt[h];
                n1 += v_[h] * v_[h];
            }
            if (!has_text) {
                o[0] = vt[0]; o[1] = vt[1];
            } else if (a.renorm_type == 0) {
                if (pass == 0) { s0 += n0; s1 += n1; continue; }
                o[0] = rbf(v_[0] * gscale); o[1] = rbf(v_[1] * gscale);                          // bf16 * 0-dim fp32 -> bf16
            } else {
                n0 = sqrtf(warp_sum(n0));
                n1 = sqrtf(warp_sum(n1));
                const float sc = fminf(fmaxf(n0 / (n1 + 1e-8f), a.renorm_min), 1.0f);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float r = __fmul_rn(v_[h], sc);                                              // bf16 * fp32[N,1] -> fp32
                    if (a.renorm_type == 2 && has_img) {
                        const int c = lane + 32 * h;
                        const float vi = c < a.C ? b2f(vi_p[base + c]) : 0.f;
                        r = __fadd_rn(vi, __fmul_rn(a.img_scale, __fsub_rn(r, vi)));             // all fp32 (bagel.py:1185)
                    }
                    o[h] = r;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = lane + 32 * h;
                if (c < a.C) a.out[(size_t)(lat0 + tok) * a.C + c] = o[h];
            }
        }
        if (!(has_text && a.renorm_type == 0)) break;
    }
}
int cfg_combine(const CfgArgs& a, int n_images, cudaStream_t s) {
    UMV_REQUIRE(a.C <= 64, UMV_ERR_UNSUPPORTED, "cfg_combine: latent dim %d > 64", a.C);
    launch_k(cfg_combine_kernel, dim3(n_images), dim3(256), 0, s, a);
    UMV_LAUNCH_CHECK("cfg_combine_kernel");
    return UMV_OK;
}

// x_t <- x_t - v * dt (bagel.py:983): a bf16-valued v makes v*dt a bf16 tensor (0-dim fp32 dt does not promote).
__global__ void euler_kernel(float* __restrict__ x, const float* __restrict__ v, int64_t n, float dt, int v_is_bf16) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float step = __fmul_rn(v[i], dt);
        x[i] = __fsub_rn(x[i], v_is_bf16 ? rbf(step) : step);
    }
}
int euler_step(float* x, const float* v, int64_t n, float dt, int v_is_bf16, cudaStream_t s) {
    if (n <= 0) return UMV_OK;
    launch_k(euler_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, x, v, n, dt, v_is_bf16);
    UMV_LAUNCH_CHECK("euler_kernel");
    return UMV_OK;
}

}  // namespace umv
