// Backward / training-step kernels of the transformer part that are HBM-bound SIMT work:
//   LayerNorm backward (+ the temporal self-subtract backward, + residual-gradient accumulation),
//   erf-GELU forward / backward, fp32 -> bf16 cast, [M, C] -> [C, M] transpose with column sums (operands of
//   the weight-gradient GEMMs and the bias gradients), head / token-build backward, AdamW.
// They back `loss.backward()` / `optimizer.step()` of the reference's training loop (train_CNN.py:513-533) for the
// modules of network/vivit/module.py and network/vivit/vivit.py; each entry point cites the forward lines.
#include "common.cuh"
#include "ptx.cuh"
#include "simt_util.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>

namespace istvt {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pk2(float a, float b) {
    const __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ void st4(bf16* p, const float (&v)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pk2(v[0], v[1]), pk2(v[2], v[3]));
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// ------------------------------------------------------------------------------------------
// LayerNorm backward, one warp per row, the row in registers (dim <= 768, dim % 4 == 0).
//   x^ = (x - mean) * rstd;  a = gamma * dy;  dx = rstd * (a - mean(a) - x^ * mean(a * x^))
//   dgamma += sum_rows dy * x^;  dbeta += sum_rows dy          (per-lane partials -> smem -> global atomics)
// dy_total = dy + dy2[row] - dy2[row + tokens] for frames 1..F-2 when dy2 != NULL: the backward of
//   res = cat(xn[:, :2], xn[:, 2:] - xn[:, 1:-1]) (module.py:192) fused in front of LN1's backward.
// ACCUM: g (fp32 residual-stream gradient) += dx, optional bf16 copy of the updated g (the next GEMM's operand).
// ------------------------------------------------------------------------------------------
constexpr int LNB_MAXC = 6;
template <typename TX, bool ACCUM>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ dy2, int frames, int tokens_pf,
                     const TX* __restrict__ x, const float* __restrict__ gamma, float* __restrict__ g,
                     bf16* __restrict__ g_bf, bf16* __restrict__ dx_out, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ out_colsum, int64_t rows, int dim, int64_t ld_dy,
                     int64_t ld_x, int64_t ld_g, int64_t ld_gb, int64_t ld_dx, float eps, int p_prefetch) {
    // dgamma / dbeta partials live in a per-warp slab of shared memory (each lane owns its columns, no atomics):
    // keeping them in registers cost 48 registers per thread and held the kernel to one CTA per SM (1.4 TB/s).
    // out_colsum (optional): column sums of the OUTPUT rows (the updated g, or dx) — the bias gradient of the nn.Linear
    // whose output gradient this kernel produces (b_2 / b_so / b_to): a third per-warp slab instead of a separate pass
    // over the rows this kernel has just written (istvt_colsum: 4 launches and 1.6 GB of re-reads per layer).
    extern __shared__ __align__(16) float s_part[];          // [warps][2 or 3][dim]
    const int lane = threadIdx.x & 31;
    const int nch = dim >> 2;
    const float inv_dim = 1.0f / static_cast<float>(dim);
    const int slabs = out_colsum != nullptr ? 3 : 2;
    float* s_dg = s_part + (threadIdx.x >> 5) * slabs * dim;
    float* s_db = s_dg + dim;
    float* s_cs = s_db + dim;                                // only touched when out_colsum != nullptr
    for (int i = lane; i < dim; i += 32) { s_dg[i] = 0.f; s_db[i] = 0.f; }
    if (out_colsum != nullptr)
        for (int i = lane; i < dim; i += 32) s_cs[i] = 0.f;
    __syncwarp();

    const int64_t wid = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t wstride = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    // Two dependent DRAM round trips per row (x / dy, then the read-modify-write of g) with one row per warp in flight
    // held the kernel at 3.4 TB/s (profiles/README.md r6f).  Each lane now pulls one 128-byte line of THIS row's g and
    // of the NEXT row's x / dy / dy2 / g into L2 before the row's own loads: no registers, the later loads hit in L2.
    auto prefetch_row = [&](int64_t r, bool with_inputs) {
        if (r >= rows) return;
        if (with_inputs) {
            if (lane * (128 / static_cast<int>(sizeof(TX))) < dim)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(x + r * ld_x + lane * (128 / sizeof(TX))));
            if (lane * 64 < dim) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(dy + r * ld_dy + lane * 64));
                if (dy2 != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(dy2 + r * ld_dy + lane * 64));
            }
        }
        if (ACCUM && lane * 32 < dim) asm volatile("prefetch.global.L2 [%0];" ::"l"(g + r * ld_g + lane * 32));
    };
    for (int64_t row = wid; row < rows; row += wstride) {
        if (p_prefetch) {
            prefetch_row(row, false);
            prefetch_row(row + wstride, true);
        }
        float xv[LNB_MAXC][4], dv[LNB_MAXC][4];
        bool sub_next = false;
        if (dy2 != nullptr) {
            const int f = static_cast<int>((row / tokens_pf) % frames);
            sub_next = f >= 1 && f <= frames - 2;
        }
        // issue every global load of the row up front (x, dy, dy2): the statistics below do not depend on dy, and
        // with one row per warp in flight the kernel is latency-bound otherwise
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nch) {
                ld4(x + row * ld_x + 4 * c, xv[i]);
                ld4(dy + row * ld_dy + 4 * c, dv[i]);
            } else {
                xv[i][0] = xv[i][1] = xv[i][2] = xv[i][3] = 0.f;
                dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
            }
        }
        if (dy2 != nullptr) {
#pragma unroll
            for (int i = 0; i < LNB_MAXC; ++i) {
                const int c = lane + 32 * i;
                if (c < nch) {
                    float t[4];
                    ld4(dy2 + row * ld_dy + 4 * c, t);
#pragma unroll
                    for (int e = 0; e < 4; ++e) dv[i][e] += t[e];
                    if (sub_next) {
                        ld4(dy2 + (row + tokens_pf) * ld_dy + 4 * c, t);
#pragma unroll
                        for (int e = 0; e < 4; ++e) dv[i][e] -= t[e];
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < LNB_MAXC; ++i) s += xv[i][0] + xv[i][1] + xv[i][2] + xv[i][3];
        const float mean = warp_sum(s) * inv_dim;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nch) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float d = xv[i][e] - mean; sq = fmaf(d, d, sq); }
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) * inv_dim + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nch) {
                float gm[4], pg[4], pb[4];
                ld4(gamma + 4 * c, gm);
                ld4(s_dg + 4 * c, pg);
                ld4(s_db + 4 * c, pb);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float xh = (xv[i][e] - mean) * rstd;
                    xv[i][e] = xh;
                    const float a = gm[e] * dv[i][e];
                    s1 += a;
                    s2 = fmaf(a, xh, s2);
                    pg[e] = fmaf(dv[i][e], xh, pg[e]);
                    pb[e] += dv[i][e];
                    dv[i][e] = a;
                }
                st4(s_dg + 4 * c, pg);
                st4(s_db + 4 * c, pb);
            }
        }
        s1 = warp_sum(s1) * inv_dim;
        s2 = warp_sum(s2) * inv_dim;
#pragma unroll
        for (int i = 0; i < LNB_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nch) {
                float dx[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) dx[e] = rstd * (dv[i][e] - s1 - xv[i][e] * s2);
                if (ACCUM) {
                    float gv[4];
                    ld4(g + row * ld_g + 4 * c, gv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) gv[e] += dx[e];
                    st4(g + row * ld_g + 4 * c, gv);
                    if (g_bf != nullptr) st4(g_bf + row * ld_gb + 4 * c, gv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) dx[e] = gv[e];       // what the column sums take below
                } else {
                    st4(dx_out + row * ld_dx + 4 * c, dx);
                }
                if (out_colsum != nullptr) {
                    float pc[4];
                    ld4(s_cs + 4 * c, pc);
#pragma unroll
                    for (int e = 0; e < 4; ++e) pc[e] += dx[e];
                    st4(s_cs + 4 * c, pc);
                }
            }
        }
    }
    __syncthreads();
    const int warps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        float a = 0.f, b = 0.f, cs = 0.f;
        for (int w = 0; w < warps; ++w) {
            a += s_part[w * slabs * dim + i];
            b += s_part[w * slabs * dim + dim + i];
            if (out_colsum != nullptr) cs += s_part[w * slabs * dim + 2 * dim + i];
        }
        atomicAdd(dgamma + i, a);
        atomicAdd(dbeta + i, b);
        if (out_colsum != nullptr) atomicAdd(out_colsum + i, cs);
    }
}

// ------------------------------------------------------------------------------------------
// exact-erf GELU (nn.GELU(), module.py:28), elementwise on bf16: forward and backward
//   gelu'(x) = 0.5 * (1 + erf(x / sqrt 2)) + x * exp(-x^2 / 2) / sqrt(2 pi)
// ------------------------------------------------------------------------------------------
// erf by Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7: fp32 noise for these uses), 2 MUFU (rcp, ex2) + FMAs —
// erff + expf made both kernels XU-bound at 3.4-3.8 TB/s.  Returns Phi(x) and phi(x) together (they share the
// exponential exp(-x^2 / 2)).
__device__ __forceinline__ void gauss_cdf_pdf(float x, float& cdf, float& pdf) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(z, 0.3275911f, 1.0f)));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));   // exp(-x^2 / 2)
    const float erf_abs = fmaf(-p, e, 1.0f);
    cdf = fmaf(0.5f, copysignf(erf_abs, x), 0.5f);
    pdf = 0.3989422804014327f * e;
}

__global__ void __launch_bounds__(256) gelu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int64_t n8) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    float v[8];
    load8(x + i * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = gelu_erf_fast(v[e]);
    store8(y + i * 8, v);
}
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, bf16* __restrict__ dx, int64_t n8) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    float v[8], d[8];
    load8(x + i * 8, v);
    load8(dy + i * 8, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float cdf, pdf;
        gauss_cdf_pdf(v[e], cdf, pdf);
        d[e] *= fmaf(v[e], pdf, cdf);
    }
    store8(dx + i * 8, d);
}

// the same for a [rows, cols] matrix with row pitches (cols % 4 == 0): thread = 4 consecutive elements of one row
__global__ void __launch_bounds__(256)
cast_f32_bf16_rows_kernel(const float* __restrict__ x, int64_t ldx, bf16* __restrict__ y, int64_t ldy, int64_t rows, int c4) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows * c4) return;
    const int64_t r = i / c4;
    const int c = static_cast<int>(i - r * c4) * 4;
    float v[4];
    ld4(x + r * ldx + c, v);
    st4(y + r * ldy + c, v);
}

// dx = dy * gelu'(x) AND colsum[c] += sum_rows dx[r, c] (the bias gradient of the Linear in front of the GELU, b_1):
// thread = 8 fixed columns, walking rows with the grid stride, sums in registers -> shared memory -> one atomic per
// column and CTA.  Replaces istvt_gelu_bwd + istvt_colsum (a second pass over the 944 MB it has just written).
__global__ void __launch_bounds__(1024)
gelu_bwd_colsum_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, bf16* __restrict__ dx,
                       float* __restrict__ colsum, int64_t rows, int cols) {
    extern __shared__ float s_gc[];                      // [cols]
    for (int i = threadIdx.x; i < cols; i += blockDim.x) s_gc[i] = 0.f;
    __syncthreads();
    const int c8 = cols >> 3;
    const int rows_per_cta = blockDim.x / c8 > 0 ? blockDim.x / c8 : 1;   // cols / 8 <= 256 is required by the host
    const int cg = threadIdx.x % c8, rl = threadIdx.x / c8;
    if (rl < rows_per_cta) {
        float s[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = 0.f;
        // four rows in flight per thread (one row at a time was latency bound: 3.2 TB/s against the flat kernel's 5.9)
        const int64_t rstep = static_cast<int64_t>(gridDim.x) * rows_per_cta;
        for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta + rl; r0 < rows; r0 += 4 * rstep) {
            uint4 xv[4], dv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t r = r0 + u * rstep;
                if (r < rows) {
                    xv[u] = *reinterpret_cast<const uint4*>(x + r * cols + cg * 8);
                    dv[u] = *reinterpret_cast<const uint4*>(dy + r * cols + cg * 8);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t r = r0 + u * rstep;
                if (r >= rows) break;
                const uint32_t xw[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, dw[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
                float d[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    d[2 * j] = __uint_as_float(dw[j] << 16) * gelu_grad_fast(__uint_as_float(xw[j] << 16));
                    d[2 * j + 1] = __uint_as_float(dw[j] & 0xffff0000u) * gelu_grad_fast(__uint_as_float(xw[j] & 0xffff0000u));
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) s[e] += d[e];
                store8(dx + r * cols + cg * 8, d);
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&s_gc[cg * 8 + e], s[e]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cols; i += blockDim.x) atomicAdd(colsum + i, s_gc[i]);
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t n4) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float v[4];
    ld4(x + i * 4, v);
    st4(y + i * 4, v);
}

// ------------------------------------------------------------------------------------------
// out[c, m] = in[m, c] (bf16), out row pitch ldo >= M (pad columns are written as zero), optional
// colsum[c] += sum_m in[m, c] (the bias gradient of the Linear whose dY is being transposed).
// 64 x 64 tiles through shared memory; both global sides move 16-byte vectors.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
transpose_colsum_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, float* __restrict__ colsum, int64_t M,
                        int C, int64_t ldo) {
    // tile2[rp][c] = (in[2rp][c], in[2rp+1][c]) as one 32-bit word: the read-out side then needs 4 conflict-free
    // 32-bit loads per 16-byte store instead of 8 two-byte loads (8-way bank conflicts in the first version).
    __shared__ uint32_t tile2[32][65];
    const int64_t m0 = static_cast<int64_t>(blockIdx.x) * 64;
    const int c0 = blockIdx.y * 64;
    const int tid = threadIdx.x;
    {
        const int rp = tid >> 3, cc = (tid & 7) * 8;
        uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
        if (c0 + cc < C) {
            if (m0 + 2 * rp < M) a = *reinterpret_cast<const uint4*>(in + (m0 + 2 * rp) * C + c0 + cc);
            if (m0 + 2 * rp + 1 < M) b = *reinterpret_cast<const uint4*>(in + (m0 + 2 * rp + 1) * C + c0 + cc);
        }
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tile2[rp][cc + 2 * j] = __byte_perm(aw[j], bw[j], 0x5410);
            tile2[rp][cc + 2 * j + 1] = __byte_perm(aw[j], bw[j], 0x7632);
        }
    }
    __syncthreads();
    if (colsum != nullptr && tid < 64 && c0 + tid < C) {
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const uint32_t w = tile2[r][tid];
            s += __uint_as_float(w << 16) + __uint_as_float(w & 0xffff0000u);
        }
        atomicAdd(colsum + c0 + tid, s);
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int idx = tid + it * 256;
        const int c = idx >> 3, mg = idx & 7;          // output row c, 8 consecutive m = row pairs 4 mg .. 4 mg + 3
        if (c0 + c < C && m0 + mg * 8 < ldo) {
            const uint4 v = make_uint4(tile2[mg * 4][c], tile2[mg * 4 + 1][c], tile2[mg * 4 + 2][c], tile2[mg * 4 + 3][c]);
            *reinterpret_cast<uint4*>(out + static_cast<int64_t>(c0 + c) * ldo + m0 + mg * 8) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Head backward (vivit.py:101 restricted to token (0,0), vivit.py:144-148): one warp per clip.
// Writes the residual-stream gradient of token (0,0) (the caller zero-fills the rest of g) and accumulates the
// gradients of transformer.norm, mlp_head LayerNorm and mlp_head Linear.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
head_bwd_kernel(const float* __restrict__ tokens, int64_t rows_per_clip, const float* __restrict__ dlogits,
                const float* __restrict__ ng, const float* __restrict__ nb, const float* __restrict__ hg,
                const float* __restrict__ hb, const float* __restrict__ hw, float* __restrict__ g,
                float* __restrict__ d_ng, float* __restrict__ d_nb, float* __restrict__ d_hg, float* __restrict__ d_hb,
                float* __restrict__ d_hw, float* __restrict__ d_hbias, int dim, float eps) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const float* x = tokens + static_cast<int64_t>(b) * rows_per_clip * dim;
    float* gx = g + static_cast<int64_t>(b) * rows_per_clip * dim;
    const float inv = 1.0f / dim;
    const float dz = dlogits[b];
    float s = 0.f;
    for (int i = lane; i < dim; i += 32) s += x[i];
    const float m1 = warp_sum(s) * inv;
    s = 0.f;
    for (int i = lane; i < dim; i += 32) { const float d = x[i] - m1; s = fmaf(d, d, s); }
    const float r1 = rsqrtf(warp_sum(s) * inv + eps);
    s = 0.f;
    for (int i = lane; i < dim; i += 32) s += (x[i] - m1) * r1 * ng[i] + nb[i];
    const float m2 = warp_sum(s) * inv;
    s = 0.f;
    for (int i = lane; i < dim; i += 32) { const float d = (x[i] - m1) * r1 * ng[i] + nb[i] - m2; s = fmaf(d, d, s); }
    const float r2 = rsqrtf(warp_sum(s) * inv + eps);
    // through mlp_head: z = sum(w * hw) + bias, w = uh * hg + hb
    float a_s1 = 0.f, a_s2 = 0.f;
    for (int i = lane; i < dim; i += 32) {
        const float u = (x[i] - m1) * r1 * ng[i] + nb[i];
        const float uh = (u - m2) * r2;
        const float w = uh * hg[i] + hb[i];
        const float dw = dz * hw[i];
        atomicAdd(d_hw + i, dz * w);
        atomicAdd(d_hg + i, dw * uh);
        atomicAdd(d_hb + i, dw);
        const float a = dw * hg[i];
        a_s1 += a;
        a_s2 = fmaf(a, uh, a_s2);
    }
    if (lane == 0) atomicAdd(d_hbias, dz);
    a_s1 = warp_sum(a_s1) * inv;
    a_s2 = warp_sum(a_s2) * inv;
    float b_s1 = 0.f, b_s2 = 0.f;
    for (int i = lane; i < dim; i += 32) {
        const float xh = (x[i] - m1) * r1;
        const float u = xh * ng[i] + nb[i];
        const float uh = (u - m2) * r2;
        const float du = r2 * (dz * hw[i] * hg[i] - a_s1 - uh * a_s2);
        atomicAdd(d_ng + i, du * xh);
        atomicAdd(d_nb + i, du);
        const float a = du * ng[i];
        b_s1 += a;
        b_s2 = fmaf(a, xh, b_s2);
    }
    b_s1 = warp_sum(b_s1) * inv;
    b_s2 = warp_sum(b_s2) * inv;
    for (int i = lane; i < dim; i += 32) {
        const float xh = (x[i] - m1) * r1;
        const float u = xh * ng[i] + nb[i];
        const float uh = (u - m2) * r2;
        const float du = r2 * (dz * hw[i] * hg[i] - a_s1 - uh * a_s2);
        gx[i] = r1 * (du * ng[i] - b_s1 - xh * b_s2);
    }
}

// ------------------------------------------------------------------------------------------
// Token-build backward (vivit.py:136-140): g is the gradient of the token buffer [B, T+1, P, D].
//   d pos_embedding[t, p, :] = sum_b g[b, t+1, p, :]          (all P positions, incl. the space token)
//   d space_token           = sum_{b, t} g[b, t+1, 0, :]
//   d temporal_token        = sum_{b, p} g[b, 0, p, :]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
token_bwd_pos_kernel(const float* __restrict__ g, float* __restrict__ dpos, float* __restrict__ dspace, int batch,
                     int t_frames, int tpf, int dim) {
    const int d4 = dim >> 2;
    const int64_t total = static_cast<int64_t>(t_frames) * tpf * d4;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = static_cast<int>(idx % d4);
    const int64_t tp = idx / d4;                 // t * tpf + p
    const int p = static_cast<int>(tp % tpf);
    const int t = static_cast<int>(tp / tpf);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < batch; ++b) {
        float v[4];
        ld4(g + ((static_cast<int64_t>(b) * (t_frames + 1) + t + 1) * tpf + p) * dim + 4 * c, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] += v[e];
    }
    float* o = dpos + tp * dim + 4 * c;
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] += acc[e];
    if (p == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) atomicAdd(dspace + 4 * c + e, acc[e]);
    }
}
__global__ void __launch_bounds__(256)
token_bwd_temporal_kernel(const float* __restrict__ g, float* __restrict__ dtemporal, int t_frames, int tpf, int dim) {
    const int b = blockIdx.x;
    const float* base = g + static_cast<int64_t>(b) * (t_frames + 1) * tpf * dim;   // frame 0 of clip b
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        float acc = 0.f;
        for (int p = 0; p < tpf; ++p) acc += base[static_cast<int64_t>(p) * dim + c];
        atomicAdd(dtemporal + c, acc);
    }
}

// ------------------------------------------------------------------------------------------
// AdamW, torch.optim.AdamW semantics (train_CNN.py:199): decoupled weight decay, bias-corrected moments.
// One launch over a flat fp32 parameter / gradient / moment buffer.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n4, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
             float grad_scale) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float pv[4], gv[4], mv[4], vv[4];
    ld4(p + 4 * i, pv); ld4(g + 4 * i, gv); ld4(m + 4 * i, mv); ld4(v + 4 * i, vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float gr = gv[e] * grad_scale;
        pv[e] *= 1.0f - lr * wd;
        mv[e] = beta1 * mv[e] + (1.0f - beta1) * gr;
        vv[e] = beta2 * vv[e] + (1.0f - beta2) * gr * gr;
        const float denom = sqrtf(vv[e]) / bc2_sqrt + eps;
        pv[e] -= (lr / bc1) * (mv[e] / denom);
    }
    st4(p + 4 * i, pv); st4(m + 4 * i, mv); st4(v + 4 * i, vv);
}

// ------------------------------------------------------------------------------------------
// Relevance rollout, one row: v[n, :] <- v[n, :] (I + C[n]) = v + v C  (C: [n, L, L], L <= 1024).
// Only row 0 (the class token's row) of the rolled-out product R = (I + C_12) ... (I + C_1) is ever read
// (visualize_rel.py:257-262), and e_0^T R can be accumulated from the last layer to the first as vector-matrix
// products — which is the order the backward pass produces the C_l in.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rollout_row_kernel(float* __restrict__ v, const float* __restrict__ cmat, int len) {
    __shared__ float sv[1024];
    const int64_t n = blockIdx.x;
    float* vr = v + n * len;
    const float* cm = cmat + n * len * len;
    for (int i = threadIdx.x; i < len; i += blockDim.x) sv[i] = vr[i];
    __syncthreads();
    for (int j = threadIdx.x; j < len; j += blockDim.x) {
        float acc = sv[j];
        for (int i = 0; i < len; ++i) acc = fmaf(sv[i], cm[static_cast<int64_t>(i) * len + j], acc);
        vr[j] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// colsum[c] += sum_m x[m, c]  (bf16 in, fp32 out): bias gradient of an nn.Linear from its output gradient.
// thread = (row lane, 8-column group), groups fastest; per-CTA partials through shared memory.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
colsum_kernel(const bf16* __restrict__ x, float* __restrict__ colsum, int64_t m, int c, int64_t ld) {
    extern __shared__ float s_cs[];
    for (int i = threadIdx.x; i < c; i += blockDim.x) s_cs[i] = 0.f;
    __syncthreads();
    const int c8 = c >> 3;
    const int lanes = blockDim.x / c8 > 0 ? blockDim.x / c8 : 1;
    const int cg = threadIdx.x % c8, rl = threadIdx.x / c8;
    if (rl < lanes) {
        float s[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = 0.f;
        for (int64_t r = static_cast<int64_t>(blockIdx.x) * lanes + rl; r < m; r += static_cast<int64_t>(gridDim.x) * lanes) {
            float v[8];
            load8(x + r * ld + cg * 8, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) s[e] += v[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&s_cs[cg * 8 + e], s[e]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(colsum + i, s_cs[i]);
}

// ------------------------------------------------------------------------------------------
// Row gather: dst[o, r, :] = src[o * outer_stride + r * row_stride + (0 .. row_bytes)], 16-byte units.
// Used by the pruned last transformer layer (only token (0,0) of every clip reaches the head, vivit.py:144-148):
// frame-0 rows of a clip are contiguous, clips are (T+1)*362 rows apart.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n_outer, int64_t outer_stride16,
                   int64_t rows, int64_t row_stride16, int row16) {
    const int64_t total = n_outer * rows * row16;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = static_cast<int>(idx % row16);
    const int64_t t = idx / row16;
    const int64_t r = t % rows;
    const int64_t o = t / rows;
    dst[idx] = src[o * outer_stride16 + r * row_stride16 + c];
}

static inline unsigned nblk(int64_t total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }

}  // namespace istvt

using namespace istvt;

extern "C" int istvt_layernorm_bwd_ld(const void* dy, const void* dy2, int64_t ld_dy, int frames, int tokens_per_frame,
                                      const void* x, int x_dtype, int64_t ld_x, const float* gamma, float* g_accum,
                                      int64_t ld_g, void* g_bf16, int64_t ld_gb, void* dx_out, int64_t ld_dx,
                                      float* dgamma, float* dbeta, float* out_colsum, int64_t rows, int dim, float eps,
                                      istvt_stream_t stream) {
    ISTVT_REQUIRE(dy && x && gamma && dgamma && dbeta && rows > 0);
    ISTVT_REQUIRE(dim % 4 == 0 && dim <= 768);
    ISTVT_REQUIRE(ld_dy >= dim && ld_x >= dim && ld_dy % 4 == 0 && ld_x % 4 == 0);
    ISTVT_REQUIRE(g_accum == nullptr || (ld_g >= dim && ld_g % 4 == 0 && (g_bf16 == nullptr || (ld_gb >= dim && ld_gb % 4 == 0))));
    ISTVT_REQUIRE(dx_out == nullptr || (ld_dx >= dim && ld_dx % 4 == 0));
    ISTVT_REQUIRE((g_accum != nullptr) != (dx_out != nullptr));
    ISTVT_REQUIRE(dy2 == nullptr || (frames > 0 && tokens_per_frame > 0));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 2;     // 2 resident CTAs per SM (46 KB smem, <= 128 registers)
    if (blocks > cap) blocks = cap;
    const bf16* d1 = static_cast<const bf16*>(dy);
    const bf16* d2 = static_cast<const bf16*>(dy2);
    bf16* gb = static_cast<bf16*>(g_bf16);
    bf16* dxo = static_cast<bf16*>(dx_out);
    const unsigned gr = static_cast<unsigned>(blocks);
    const size_t smem = 8 * (out_colsum != nullptr ? 3 : 2) * static_cast<size_t>(dim) * sizeof(float);
    // ISTVT_LNB_PREFETCH=0: without the L2 prefetches of the next row (A/B measurements)
    static const int pf = []() { const char* e = getenv("ISTVT_LNB_PREFETCH"); return e ? atoi(e) : 1; }();
    ISTVT_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 768 * 4));
    ISTVT_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 768 * 4));
    ISTVT_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<bf16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 768 * 4));
    ISTVT_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<bf16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 768 * 4));
    if (x_dtype == ISTVT_F32) {
        const float* xx = static_cast<const float*>(x);
        if (g_accum)
            layernorm_bwd_kernel<float, true><<<gr, 256, smem, st>>>(d1, d2, frames, tokens_per_frame, xx, gamma, g_accum, gb,
                                                                   nullptr, dgamma, dbeta, out_colsum, rows, dim, ld_dy, ld_x, ld_g, ld_gb, ld_dx, eps, pf);
        else
            layernorm_bwd_kernel<float, false><<<gr, 256, smem, st>>>(d1, d2, frames, tokens_per_frame, xx, gamma, nullptr,
                                                                    nullptr, dxo, dgamma, dbeta, out_colsum, rows, dim, ld_dy, ld_x, ld_g, ld_gb, ld_dx, eps, pf);
    } else if (x_dtype == ISTVT_BF16) {
        const bf16* xx = static_cast<const bf16*>(x);
        if (g_accum)
            layernorm_bwd_kernel<bf16, true><<<gr, 256, smem, st>>>(d1, d2, frames, tokens_per_frame, xx, gamma, g_accum, gb,
                                                                  nullptr, dgamma, dbeta, out_colsum, rows, dim, ld_dy, ld_x, ld_g, ld_gb, ld_dx, eps, pf);
        else
            layernorm_bwd_kernel<bf16, false><<<gr, 256, smem, st>>>(d1, d2, frames, tokens_per_frame, xx, gamma, nullptr,
                                                                   nullptr, dxo, dgamma, dbeta, out_colsum, rows, dim, ld_dy, ld_x, ld_g, ld_gb, ld_dx, eps, pf);
    } else {
        return ISTVT_ERR_INVALID_ARG;
    }
    count_launch();
    return launch_status();
}

extern "C" int istvt_layernorm_bwd(const void* dy, const void* dy2, int frames, int tokens_per_frame, const void* x,
                                   int x_dtype, const float* gamma, float* g_accum, void* g_bf16, void* dx_out,
                                   float* dgamma, float* dbeta, int64_t rows, int dim, float eps,
                                   istvt_stream_t stream) {
    return istvt_layernorm_bwd_ld(dy, dy2, dim, frames, tokens_per_frame, x, x_dtype, dim, gamma, g_accum, dim, g_bf16, dim,
                                  dx_out, dim, dgamma, dbeta, nullptr, rows, dim, eps, stream);
}

extern "C" int istvt_gelu_fwd(const void* x, void* y, int64_t n, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && y && n > 0 && n % 8 == 0);
    gelu_fwd_kernel<<<nblk(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(x), static_cast<bf16*>(y), n / 8);
    count_launch();
    return launch_status();
}

extern "C" int istvt_gelu_bwd(const void* dy, const void* x, void* dx, int64_t n, istvt_stream_t stream) {
    ISTVT_REQUIRE(dy && x && dx && n > 0 && n % 8 == 0);
    gelu_bwd_kernel<<<nblk(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(dy), static_cast<const bf16*>(x), static_cast<bf16*>(dx), n / 8);
    count_launch();
    return launch_status();
}

extern "C" int istvt_gelu_bwd_colsum(const void* dy, const void* x, void* dx, float* colsum, int64_t rows, int cols,
                                     istvt_stream_t stream) {
    ISTVT_REQUIRE(dy && x && dx && colsum && rows > 0 && cols > 0 && cols % 8 == 0);
    const int c8 = cols / 8;
    // thread = 8 columns: up to 1024 threads per CTA so that one CTA covers whole rows (cols <= 8192)
    ISTVT_REQUIRE(c8 <= 1024);
    int threads = (c8 <= 256) ? (256 / c8) * c8 : c8;
    threads = (threads + 31) / 32 * 32;
    int64_t blocks = static_cast<int64_t>(sm_count()) * (threads <= 512 ? 4 : 2);
    ISTVT_CHECK_CUDA(cudaFuncSetAttribute(gelu_bwd_colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(cols * sizeof(float))));
    gelu_bwd_colsum_kernel<<<static_cast<unsigned>(blocks), threads, cols * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(dy), static_cast<const bf16*>(x), static_cast<bf16*>(dx), colsum, rows, cols);
    count_launch();
    return launch_status();
}

extern "C" int istvt_cast_f32_bf16(const float* x, void* y, int64_t n, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && y && n > 0 && n % 4 == 0);
    cast_f32_bf16_kernel<<<nblk(n / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<bf16*>(y), n / 4);
    count_launch();
    return launch_status();
}

extern "C" int istvt_transpose_colsum(const void* in, void* out, float* colsum, int64_t m, int c, int64_t ldo,
                                      istvt_stream_t stream) {
    ISTVT_REQUIRE(in && out && m > 0 && c > 0 && c % 8 == 0 && ldo % 8 == 0 && ldo >= m);
    const dim3 grid(static_cast<unsigned>((ldo + 63) / 64), static_cast<unsigned>((c + 63) / 64));
    transpose_colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(in), static_cast<bf16*>(out), colsum, m, c, ldo);
    count_launch();
    return launch_status();
}

extern "C" int istvt_head_bwd(const float* tokens, int64_t rows_per_clip, const float* dlogits, const float* norm_g,
                              const float* norm_b, const float* head_g, const float* head_b, const float* head_w,
                              float* g, float* d_norm_g, float* d_norm_b, float* d_head_g, float* d_head_b,
                              float* d_head_w, float* d_head_bias, int batch, int dim, float eps,
                              istvt_stream_t stream) {
    ISTVT_REQUIRE(tokens && dlogits && g && batch > 0 && dim > 0);
    head_bwd_kernel<<<batch, 32, 0, static_cast<cudaStream_t>(stream)>>>(
        tokens, rows_per_clip, dlogits, norm_g, norm_b, head_g, head_b, head_w, g, d_norm_g, d_norm_b, d_head_g,
        d_head_b, d_head_w, d_head_bias, dim, eps);
    count_launch();
    return launch_status();
}

extern "C" int istvt_token_bwd(const float* g, float* d_pos, float* d_space, float* d_temporal, int batch, int t,
                               int tokens_per_frame, int dim, istvt_stream_t stream) {
    ISTVT_REQUIRE(g && d_pos && d_space && d_temporal && batch > 0 && t > 0 && dim % 4 == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t total = static_cast<int64_t>(t) * tokens_per_frame * (dim / 4);
    token_bwd_pos_kernel<<<nblk(total, 256), 256, 0, st>>>(g, d_pos, d_space, batch, t, tokens_per_frame, dim);
    token_bwd_temporal_kernel<<<batch, 256, 0, st>>>(g, d_temporal, t, tokens_per_frame, dim);
    count_launch();
    count_launch();
    return launch_status();
}

extern "C" int istvt_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                float grad_scale, istvt_stream_t stream) {
    ISTVT_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0 && n % 4 == 0 && step >= 1);
    const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
    const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
    adamw_kernel<<<nblk(n / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        params, grads, exp_avg, exp_avg_sq, n / 4, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale);
    count_launch();
    return launch_status();
}

extern "C" int istvt_rollout_row(float* v, const float* cmat, int64_t n, int len, istvt_stream_t stream) {
    ISTVT_REQUIRE(v && cmat && n > 0 && len > 0 && len <= 1024 && n < (int64_t(1) << 31));
    rollout_row_kernel<<<static_cast<unsigned>(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(v, cmat, len);
    count_launch();
    return launch_status();
}

// dst [n_outer, rows, row_bytes] (contiguous) <- src; strides and row_bytes in BYTES, all multiples of 16.
extern "C" int istvt_gather_rows(const void* src, void* dst, int64_t n_outer, int64_t outer_stride_bytes, int64_t rows,
                                 int64_t row_stride_bytes, int64_t row_bytes, istvt_stream_t stream) {
    ISTVT_REQUIRE(src && dst && n_outer > 0 && rows > 0 && row_bytes > 0);
    ISTVT_REQUIRE(outer_stride_bytes % 16 == 0 && row_stride_bytes % 16 == 0 && row_bytes % 16 == 0);
    ISTVT_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0);
    const int64_t total = n_outer * rows * (row_bytes / 16);
    gather_rows_kernel<<<nblk(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(src), static_cast<uint4*>(dst), n_outer, outer_stride_bytes / 16, rows,
        row_stride_bytes / 16, static_cast<int>(row_bytes / 16));
    count_launch();
    return launch_status();
}

extern "C" int istvt_colsum_ld(const void* x, int64_t ld, float* colsum, int64_t m, int c, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && colsum && m > 0 && c > 0 && c % 8 == 0 && c <= 8192 && ld >= c && ld % 8 == 0);
    const int c8 = c / 8;
    int lanes = 256 / c8;
    if (lanes < 1) lanes = 1;
    int threads = lanes * c8;
    threads = (threads + 31) / 32 * 32;
    ISTVT_REQUIRE(threads <= 1024);
    int64_t blocks = (m + 63) / 64;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    colsum_kernel<<<static_cast<unsigned>(blocks), threads, c * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(x), colsum, m, c, ld);
    count_launch();
    return launch_status();
}

extern "C" int istvt_colsum(const void* x, float* colsum, int64_t m, int c, istvt_stream_t stream) {
    return istvt_colsum_ld(x, c, colsum, m, c, stream);
}

extern "C" int istvt_cast_f32_bf16_rows(const float* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int cols,
                                        istvt_stream_t stream) {
    ISTVT_REQUIRE(x && y && rows > 0 && cols > 0 && cols % 4 == 0 && ldx >= cols && ldy >= cols && ldx % 4 == 0 && ldy % 4 == 0);
    const int64_t n4 = rows * (cols / 4);
    cast_f32_bf16_rows_kernel<<<nblk(n4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ldx, static_cast<bf16*>(y), ldy, rows, cols / 4);
    count_launch();
    return launch_status();
}
