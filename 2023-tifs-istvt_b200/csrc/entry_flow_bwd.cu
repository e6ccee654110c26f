// Training-mode kernels of the Xception entry flow (network/xception.py:52-101,193-206), NHWC bf16 activations:
//   BatchNorm2d with batch statistics: per-channel sum / sum-of-squares, finalize (+ running-stat update),
//     apply (+ReLU); backward reduce (dgamma, dbeta) and apply (dx)                    (xception.py:58,69,75,119,123)
//   MaxPool2d(3,2,1) + skip add that records the arg-max, and its backward (gather)    (xception.py:87-88,100)
//   depthwise 3x3 weight gradient (the data gradient is the forward kernel with flipped taps)   (xception.py:43,47)
//   block-input gradient: ReLU mask of the main branch + scatter of the stride-2 skip branch  (xception.py:82-85,94)
//   token-gradient gather (vivit.py:133-138 backward), im2col^T operands for the conv1 / conv2 weight gradients.
// All are HBM-bound SIMT kernels; the GEMM-shaped gradients run on the tcgen05 GEMMs (gemm_tcgen05*.cu).
#include "common.cuh"
#include "ptx.cuh"
#include "simt_util.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>

namespace istvt {

typedef __nv_bfloat16 bf16;

// thread layout shared by the per-channel reductions: thread = (row lane, 8-channel group), groups fastest
struct ChanLayout {
    int c8, lanes, cg, rl;
    bool active;
    __device__ ChanLayout(int c) {
        c8 = c >> 3;
        lanes = blockDim.x / c8;
        if (lanes < 1) lanes = 1;
        cg = threadIdx.x % c8;
        rl = threadIdx.x / c8;
        active = rl < lanes;
    }
};

// ------------------------------------------------------------------------------------------
// BatchNorm statistics: sum[c] += sum_m x[m, c], sumsq[c] += sum_m x[m, c]^2
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_stats_kernel(const bf16* __restrict__ x, float* __restrict__ sum, float* __restrict__ sumsq, int64_t m, int c) {
    extern __shared__ float s_acc[];   // [2][c]
    for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const ChanLayout L(c);
    if (L.active) {
        float s[8], q[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
        for (int64_t r = static_cast<int64_t>(blockIdx.x) * L.lanes + L.rl; r < m;
             r += static_cast<int64_t>(gridDim.x) * L.lanes) {
            float v[8];
            load8(x + r * c + L.cg * 8, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) { s[e] += v[e]; q[e] = fmaf(v[e], v[e], q[e]); }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            atomicAdd(&s_acc[L.cg * 8 + e], s[e]);
            atomicAdd(&s_acc[c + L.cg * 8 + e], q[e]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        atomicAdd(sum + i, s_acc[i]);
        atomicAdd(sumsq + i, s_acc[c + i]);
    }
}

// mean / biased var -> scale = gamma * rstd, shift = beta - mean * scale; running stats as nn.BatchNorm2d
// (momentum 0.1, running_var with the unbiased batch variance).
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int64_t m, int c, float eps, float momentum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const double mean = static_cast<double>(sum[i]) / static_cast<double>(m);
    double var = static_cast<double>(sumsq[i]) / static_cast<double>(m) - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float sc = gamma[i] * rstd;
    scale[i] = sc;
    shift[i] = beta[i] - static_cast<float>(mean) * sc;
    mean_out[i] = static_cast<float>(mean);
    rstd_out[i] = rstd;
    if (running_mean != nullptr) {
        const double unbiased = m > 1 ? var * static_cast<double>(m) / static_cast<double>(m - 1) : var;
        running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * static_cast<float>(mean);
        running_var[i] = (1.f - momentum) * running_var[i] + momentum * static_cast<float>(unbiased);
    }
}

__global__ void __launch_bounds__(256)
bn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                bf16* __restrict__ y, int64_t n8, int c8, int relu) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const int cg = static_cast<int>(i % c8);
    float v[8], sc[8], sh[8];
    load8(x + i * 8, v);
    load8(scale + cg * 8, sc);
    load8(shift + cg * 8, sh);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        v[e] = fmaf(v[e], sc[e], sh[e]);
        if (relu) v[e] = fmaxf(v[e], 0.f);
    }
    store8(y + i * 8, v);
}

// dz = dy * (relu ? y > 0 : 1) with y = x * scale + shift;  dbeta += sum dz;  dgamma += sum dz * x^
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t m, int c, int relu) {
    extern __shared__ float s_acc[];   // [2][c]
    for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const ChanLayout L(c);
    if (L.active) {
        float sc[8], sh[8], mu[8], rs[8], dg[8], db[8];
        load8(scale + L.cg * 8, sc); load8(shift + L.cg * 8, sh);
        load8(mean + L.cg * 8, mu);  load8(rstd + L.cg * 8, rs);
#pragma unroll
        for (int e = 0; e < 8; ++e) { dg[e] = 0.f; db[e] = 0.f; }
        for (int64_t r = static_cast<int64_t>(blockIdx.x) * L.lanes + L.rl; r < m;
             r += static_cast<int64_t>(gridDim.x) * L.lanes) {
            float v[8], d[8];
            load8(x + r * c + L.cg * 8, v);
            load8(dy + r * c + L.cg * 8, d);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float dz = (relu && fmaf(v[e], sc[e], sh[e]) <= 0.f) ? 0.f : d[e];
                db[e] += dz;
                dg[e] = fmaf(dz, (v[e] - mu[e]) * rs[e], dg[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            atomicAdd(&s_acc[L.cg * 8 + e], dg[e]);
            atomicAdd(&s_acc[c + L.cg * 8 + e], db[e]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        atomicAdd(dgamma + i, s_acc[i]);
        atomicAdd(dbeta + i, s_acc[c + i]);
    }
}

// dx, row-loop form: thread = (row lane, 8-channel group) with the per-channel coefficients in registers, two rows in
// flight.  dx = sc*dz + B*x + C with B = -sc*dgamma*rstd/m, C = -sc*dbeta/m - B*mean (the same expression as below,
// regrouped).  The flat form below re-loaded six 32-byte parameter vectors per 16-byte data vector: 12 of its 15
// load instructions were parameters, and it ran at ~3 TB/s.
__global__ void __launch_bounds__(256)
bn_bwd_apply_rows_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ dgamma, const float* __restrict__ dbeta, bf16* __restrict__ dx,
                         int64_t m, int c, float inv_m, int relu) {
    const ChanLayout L(c);
    if (!L.active) return;
    float sc[8], sh[8], cb[8], cc[8];
    {
        float mu[8], rs[8], dg[8], db[8];
        load8(scale + L.cg * 8, sc); load8(shift + L.cg * 8, sh);
        load8(mean + L.cg * 8, mu);  load8(rstd + L.cg * 8, rs);
        load8(dgamma + L.cg * 8, dg); load8(dbeta + L.cg * 8, db);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            cb[e] = -sc[e] * dg[e] * rs[e] * inv_m;
            cc[e] = -sc[e] * db[e] * inv_m - cb[e] * mu[e];
        }
    }
    const int64_t stride = static_cast<int64_t>(gridDim.x) * L.lanes;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * L.lanes + L.rl; r < m; r += 2 * stride) {
        const int64_t r1 = r + stride;
        const bool two = r1 < m;
        float v0[8], d0[8], v1[8], d1[8];
        load8(x + r * c + L.cg * 8, v0);
        load8(dy + r * c + L.cg * 8, d0);
        if (two) {
            load8(x + r1 * c + L.cg * 8, v1);
            load8(dy + r1 * c + L.cg * 8, d1);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float dz = (relu && fmaf(v0[e], sc[e], sh[e]) <= 0.f) ? 0.f : d0[e];
            d0[e] = fmaf(sc[e], dz, fmaf(cb[e], v0[e], cc[e]));
        }
        store8(dx + r * c + L.cg * 8, d0);
        if (two) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float dz = (relu && fmaf(v1[e], sc[e], sh[e]) <= 0.f) ? 0.f : d1[e];
                d1[e] = fmaf(sc[e], dz, fmaf(cb[e], v1[e], cc[e]));
            }
            store8(dx + r1 * c + L.cg * 8, d1);
        }
    }
}

// dx = gamma * rstd * (dz - dbeta / m - x^ * dgamma / m)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, bf16* __restrict__ dx,
                    int64_t n8, int c8, float inv_m, int relu) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const int cg = static_cast<int>(i % c8);
    float v[8], d[8], sc[8], sh[8], mu[8], rs[8], dg[8], db[8];
    load8(x + i * 8, v); load8(dy + i * 8, d);
    load8(scale + cg * 8, sc); load8(shift + cg * 8, sh);
    load8(mean + cg * 8, mu);  load8(rstd + cg * 8, rs);
    load8(dgamma + cg * 8, dg); load8(dbeta + cg * 8, db);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float dz = (relu && fmaf(v[e], sc[e], sh[e]) <= 0.f) ? 0.f : d[e];
        const float xh = (v[e] - mu[e]) * rs[e];
        d[e] = sc[e] * (dz - db[e] * inv_m - xh * dg[e] * inv_m);
    }
    store8(dx + i * 8, d);
}

// ------------------------------------------------------------------------------------------
// maxpool 3x3 s2 p1 + skip, recording the arg-max tap (ky*3+kx, first maximum in scan order as torch does).
// TOKENS: writes fp32 tokens[b, f+1, 1+p, :] + pos_emb instead of y.
// ------------------------------------------------------------------------------------------
template <bool TOKENS>
__global__ void __launch_bounds__(256)
pool_add_idx_kernel(const bf16* __restrict__ x, const bf16* __restrict__ skip, bf16* __restrict__ y,
                    const float* __restrict__ pos_emb, float* __restrict__ tokens, uint8_t* __restrict__ idx_out, int n,
                    int h, int w, int c, int ho, int wo, int t_frames) {
    const int c8 = c >> 3;
    const int64_t total = static_cast<int64_t>(n) * ho * wo * c8;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cg = static_cast<int>(idx % c8);
    int64_t t = idx / c8;
    const int ox = static_cast<int>(t % wo);
    t /= wo;
    const int oy = static_cast<int>(t % ho);
    const int img = static_cast<int>(t / ho);
    const int ch = cg * 8;
    float m[8];
    int am[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { m[e] = -INFINITY; am[e] = 0; }
    const bf16* xin = x + static_cast<int64_t>(img) * h * w * c + ch;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int iy = 2 * oy - 1 + dy;
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int ix = 2 * ox - 1 + dx;
            if (ix < 0 || ix >= w) continue;
            float v[8];
            load8(xin + (static_cast<int64_t>(iy) * w + ix) * c, v);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (v[e] > m[e]) { m[e] = v[e]; am[e] = dy * 3 + dx; }
        }
    }
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) { lo |= static_cast<uint32_t>(am[e]) << (8 * e); hi |= static_cast<uint32_t>(am[4 + e]) << (8 * e); }
    *reinterpret_cast<uint2*>(idx_out + idx * 8) = make_uint2(lo, hi);
    float s[8];
    load8(skip + idx * 8, s);
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] += s[e];
    if (!TOKENS) {
        store8(y + idx * 8, m);
    } else {
        const int tpf = ho * wo + 1;
        const int b = img / t_frames, f = img - b * t_frames;
        const int p = oy * wo + ox;
        float pe[8];
        load8(pos_emb + (static_cast<int64_t>(f) * tpf + 1 + p) * c + ch, pe);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] += pe[e];
        store8(tokens + ((static_cast<int64_t>(b) * (t_frames + 1) + f + 1) * tpf + 1 + p) * c + ch, m);
    }
}

// dx[n, iy, ix, :] = sum over the (<= 4) windows that contain (iy, ix) of dy[window] where the window's arg-max is
// this pixel.  Gather form: deterministic, no atomics.  Thread = a 2 x 2 block of input pixels x 8 channels: the block
// (2a.., 2b..) lies in the four windows (a | a+1, b | b+1) only, so 4 window loads serve 9 (pixel, window) pairs — one
// thread per pixel read 9 windows for the same four pixels (1.5 TB/s, profiles/README.md r7j); per pixel the windows are
// added in the same order as before (bit-identical).
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ amax, bf16* __restrict__ dx, int n, int h, int w,
                int c, int ho, int wo) {
    // grid = (ceil(ceil(w/2) * c/8 / 256), ceil(h/2), n): one 32-bit division per thread
    const int c8 = c >> 3;
    const int wp = (w + 1) >> 1;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;        // b * c8 + cg
    if (col >= wp * c8) return;
    const int b = col / c8;
    const int cg = col - b * c8;
    const int a = blockIdx.y;
    const int img = blockIdx.z;
    float acc[2][2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[i][j][e] = 0.f;
#pragma unroll
    for (int wy = 0; wy < 2; ++wy) {
        const int oy = a + wy;
        if (oy >= ho) continue;
#pragma unroll
        for (int wx = 0; wx < 2; ++wx) {
            const int ox = b + wx;
            if (ox >= wo) continue;
            const int64_t o = ((static_cast<int64_t>(img) * ho + oy) * wo + ox) * c + cg * 8;
            const uint2 am = *reinterpret_cast<const uint2*>(amax + o);
            float d[8];
            load8(dy + o, d);
            // pixel (2a + py, 2b + px) sits at (ky, kx) = (py - 2 wy + 1, px - 2 wx + 1) of this window
#pragma unroll
            for (int py = wy; py < 2; ++py) {
#pragma unroll
                for (int px = wx; px < 2; ++px) {
                    const int code = (py - 2 * wy + 1) * 3 + (px - 2 * wx + 1);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const uint32_t word = e < 4 ? am.x : am.y;
                        if (static_cast<int>((word >> (8 * (e & 3))) & 0xffu) == code) acc[py][px][e] += d[e];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int py = 0; py < 2; ++py) {
        const int iy = 2 * a + py;
        if (iy >= h) continue;
#pragma unroll
        for (int px = 0; px < 2; ++px) {
            const int ix = 2 * b + px;
            if (ix >= w) continue;
            store8(dx + ((static_cast<int64_t>(img) * h + iy) * w + ix) * c + cg * 8, acc[py][px]);
        }
    }
}

// d_out[n, oy, ox, :] (bf16) = g[b, f+1, 1 + oy*wo + ox, :] (fp32 token gradient)
__global__ void __launch_bounds__(256)
token_grad_gather_kernel(const float* __restrict__ g, bf16* __restrict__ d_out, int n, int t_frames, int tpf, int c) {
    const int c8 = c >> 3;
    const int64_t total = static_cast<int64_t>(n) * (tpf - 1) * c8;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cg = static_cast<int>(idx % c8);
    int64_t t = idx / c8;
    const int p = static_cast<int>(t % (tpf - 1));
    const int img = static_cast<int>(t / (tpf - 1));
    const int b = img / t_frames, f = img - b * t_frames;
    float v[8];
    load8(g + ((static_cast<int64_t>(b) * (t_frames + 1) + f + 1) * tpf + 1 + p) * c + cg * 8, v);
    store8(d_out + idx * 8, v);
}

// ------------------------------------------------------------------------------------------
// depthwise 3x3 weight gradient: dw[ky][kx][c] += sum_{n,y,x} in[n, y+ky-1, x+kx-1, c] * dy[n, y, x, c]
// (in = relu(x) when the forward applied the ReLU on load).
// Same machinery as the forward kernel (entry_flow.cu): persistent CTAs, per item ONE TMA load of the x halo tile
// (10 x 18 x 64 channels; pad-1 border = TMA zero fill) and one of the dy tile (8 x 16 x 64) into a double-buffered
// slot; thread = (4 channels, one tile column) walks down the rows with the last three dy rows in registers and
// 9 x 4 accumulators.  Items are ordered channel-group-major and every CTA takes a contiguous range, so the
// accumulators live in registers across tiles and are flushed (shared atomics -> global atomics) only when the
// channel group changes: the first version (one CTA per strip, 9 uncached neighbour loads per pixel) ran at 0.8 TB/s.
// ------------------------------------------------------------------------------------------
constexpr int DWG_TW = 16, DWG_TH = 8, DWG_CG = 64, DWG_THREADS = 256;
constexpr int DWG_X_BYTES = (DWG_TH + 2) * (DWG_TW + 2) * DWG_CG * 2;     // 23040
constexpr int DWG_DY_BYTES = DWG_TH * DWG_TW * DWG_CG * 2;                // 16384
constexpr int DWG_STAGE = DWG_X_BYTES + DWG_DY_BYTES;                     // 39424 (multiple of 128)
constexpr int DWG_SMEM = 2 * DWG_STAGE + 9 * DWG_CG * 4 + 128 + 64;

__device__ __forceinline__ void dwg_lds4(const bf16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}

__global__ void __launch_bounds__(DWG_THREADS, 2)
dwconv_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy,
                        float* __restrict__ dw, int n, int h, int w, int c, int relu_in) {
    extern __shared__ __align__(128) uint8_t dwg_smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dwg_smem) + 127) & ~uintptr_t(127));
    float* s_acc = reinterpret_cast<float*>(base + 2 * DWG_STAGE);              // [9][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * DWG_STAGE + 9 * DWG_CG * 4);

    const int tiles_x = (w + DWG_TW - 1) / DWG_TW;
    const int tiles_y = (h + DWG_TH - 1) / DWG_TH;
    const int cgroups = (c + DWG_CG - 1) / DWG_CG;
    const int64_t per_cg = static_cast<int64_t>(n) * tiles_y * tiles_x;
    const int64_t total = per_cg * cgroups;
    const int64_t per_cta = (total + gridDim.x - 1) / gridDim.x;
    const int64_t it0 = per_cta * blockIdx.x;
    const int64_t it1 = (it0 + per_cta < total) ? it0 + per_cta : total;

    const int tid = threadIdx.x;
    const int cq = tid & 15;          // channel quad inside the 64-channel group
    const int col = tid >> 4;         // tile column
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_dy);
    }
    for (int i = tid; i < 9 * DWG_CG; i += DWG_THREADS) s_acc[i] = 0.f;
    __syncthreads();

    auto decode = [&](int64_t item, int& tx, int& ty, int& img, int& cg) {
        cg = static_cast<int>(item / per_cg);
        int64_t r = item - cg * per_cg;
        tx = static_cast<int>(r % tiles_x);
        r /= tiles_x;
        ty = static_cast<int>(r % tiles_y);
        img = static_cast<int>(r / tiles_y);
    };
    auto issue = [&](int64_t item, int buf) {
        int tx, ty, img, cg;
        decode(item, tx, ty, img, cg);
        mbar_arrive_expect_tx(&bars[buf], DWG_STAGE);
        tma_load_4d(base + buf * DWG_STAGE, &tm_x, &bars[buf], cg * DWG_CG, tx * DWG_TW - 1, ty * DWG_TH - 1, img);
        tma_load_4d(base + buf * DWG_STAGE + DWG_X_BYTES, &tm_dy, &bars[buf], cg * DWG_CG, tx * DWG_TW, ty * DWG_TH, img);
    };

    float acc[9][4];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[k][e] = 0.f;
    int cg_cur = -1;

    auto flush = [&](int cg) {   // all threads: fold the register partials of channel group cg into dw
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                atomicAdd(&s_acc[k * DWG_CG + cq * 4 + e], acc[k][e]);
                acc[k][e] = 0.f;
            }
        __syncthreads();
        for (int i = tid; i < 9 * DWG_CG; i += DWG_THREADS) {
            const int k = i / DWG_CG, ch = cg * DWG_CG + (i - k * DWG_CG);
            if (ch < c) atomicAdd(dw + k * c + ch, s_acc[i]);
            s_acc[i] = 0.f;
        }
        __syncthreads();
    };

    if (tid == 0 && it0 < it1) issue(it0, 0);
    int buf = 0;
    uint32_t phase[2] = {0, 0};
    for (int64_t item = it0; item < it1; ++item) {
        if (tid == 0 && item + 1 < it1) issue(item + 1, buf ^ 1);   // slot buf^1 was released by the barrier below
        int tx, ty, img, cg;
        decode(item, tx, ty, img, cg);
        if (cg != cg_cur) {
            if (cg_cur >= 0) flush(cg_cur);
            cg_cur = cg;
        }
        mbar_wait(&bars[buf], phase[buf]);
        phase[buf] ^= 1;
        const bf16* xt = reinterpret_cast<const bf16*>(base + buf * DWG_STAGE) + col * DWG_CG + cq * 4;
        const bf16* dt = reinterpret_cast<const bf16*>(base + buf * DWG_STAGE + DWG_X_BYTES) + col * DWG_CG + cq * 4;
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4];   // dy rows i-2, i-1, i
#pragma unroll
        for (int i = 0; i < DWG_TH + 2; ++i) {     // input (halo) row i pairs with dy rows i, i-1, i-2 under ky = 0, 1, 2
            if (i < DWG_TH) dwg_lds4(dt + i * DWG_TW * DWG_CG, d2);
            else d2[0] = d2[1] = d2[2] = d2[3] = 0.f;
            float v[3][4];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                dwg_lds4(xt + (i * (DWG_TW + 2) + kx) * DWG_CG, v[kx]);
                if (relu_in) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[kx][e] = fmaxf(v[kx][e], 0.f);
                }
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc[0 * 3 + kx][e] = fmaf(v[kx][e], d2[e], acc[0 * 3 + kx][e]);
                    acc[1 * 3 + kx][e] = fmaf(v[kx][e], d1[e], acc[1 * 3 + kx][e]);
                    acc[2 * 3 + kx][e] = fmaf(v[kx][e], d0[e], acc[2 * 3 + kx][e]);
                }
#pragma unroll
            for (int e = 0; e < 4; ++e) { d0[e] = d1[e]; d1[e] = d2[e]; }
        }
        __syncthreads();   // everyone is done reading slot `buf`
        buf ^= 1;
    }
    if (cg_cur >= 0) flush(cg_cur);
}

// dx_in = d_main * (relu_in ? x_in > 0 : 1) + (even pixel ? d_skip[n, y/2, x/2, :] : 0)
__global__ void __launch_bounds__(256)
block_input_grad_kernel(const bf16* __restrict__ d_main, const bf16* __restrict__ x_in, const bf16* __restrict__ d_skip,
                        bf16* __restrict__ dx, int n, int h, int w, int c, int relu_in) {
    const int c8 = c >> 3;
    const int64_t total = static_cast<int64_t>(n) * h * w * c8;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cg = static_cast<int>(idx % c8);
    int64_t t = idx / c8;
    const int ix = static_cast<int>(t % w);
    t /= w;
    const int iy = static_cast<int>(t % h);
    const int img = static_cast<int>(t / h);
    float d[8];
    load8(d_main + idx * 8, d);
    if (relu_in) {
        float v[8];
        load8(x_in + idx * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) if (v[e] <= 0.f) d[e] = 0.f;
    }
    if (((iy | ix) & 1) == 0) {
        const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
        float s[8];
        load8(d_skip + ((static_cast<int64_t>(img) * ho + (iy >> 1)) * wo + (ix >> 1)) * c + cg * 8, s);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] += s[e];
    }
    store8(dx + idx * 8, d);
}

// ------------------------------------------------------------------------------------------
// im2col^T operands of the dense-convolution weight gradients (K-major over the output pixels m):
//   conv2 (3x3 s1 p0, NHWC bf16 input [n, h, w, cin]):  out[(ky*3+kx)*cin + ci, m] = x[n, oy+ky, ox+kx, ci]
//   conv1 (3x3 s2 p0, NCHW fp32 input [n, 3, h, w])  :  out[ci*9 + ky*3 + kx, m]  = x[n, ci, 2oy+ky, 2ox+kx]
//                                                       (rows 27..31 zero so that K' = 32)
// ------------------------------------------------------------------------------------------
// CTA = 256 consecutive output pixels x all 9 taps.  The pixel -> input offset map is computed once; per tap every thread
// loads 16 B (8 channels) of two neighbouring pixels and writes them TRANSPOSED into shared memory as 8 packed 32-bit
// words (conflict-free: consecutive lanes = consecutive pixel pairs), then the tile leaves as 512-byte row segments with
// 16-byte stores.  (The first version — 64 pixels x 1 tap per CTA, transposition on the read side of shared memory with
// 8-way bank conflicts, three 64-bit divisions per load — wrote at 1.0 TB/s, profiles/README.md r7j.)
constexpr int I2C_PIX = 256;
__global__ void __launch_bounds__(256)
im2col_t_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int n, int h, int w, int cin, int64_t ldo) {
    __shared__ int64_t base[I2C_PIX];                       // input pixel index of (img, oy, ox), -1 past the end
    __shared__ __align__(16) uint32_t tile_t[64][I2C_PIX / 2 + 4];   // [channel][pixel pair], row pitch 528 B (16-byte aligned)
    const int ho = h - 2, wo = w - 2;
    const int64_t m_total = static_cast<int64_t>(n) * ho * wo;
    const int64_t m0 = static_cast<int64_t>(blockIdx.x) * I2C_PIX;
    {
        const int64_t m = m0 + threadIdx.x;
        int64_t bidx = -1;
        if (m < m_total) {
            const int ox = static_cast<int>(m % wo);
            const int64_t q = m / wo;
            const int oy = static_cast<int>(q % ho);
            const int64_t img = q / ho;
            bidx = (img * h + oy) * w + ox;
        }
        base[threadIdx.x] = bidx;
    }
    __syncthreads();
    const int chunks = cin >> 3;                            // 16-byte channel chunks per pixel
    const int n_ld = (I2C_PIX / 2) * chunks;                // (pixel pair, chunk) loads per tap
    const int n_st = cin * (I2C_PIX / 8);                   // 16-byte stores per tap
    for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - ky * 3;
        const int64_t shift = static_cast<int64_t>(ky) * w + kx;
        for (int i = threadIdx.x; i < n_ld; i += blockDim.x) {
            const int pr = i % (I2C_PIX / 2), cc = (i / (I2C_PIX / 2)) * 8;     // consecutive lanes = consecutive pairs
            const int64_t b0 = base[2 * pr], b1 = base[2 * pr + 1];
            uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
            if (b0 >= 0) v0 = *reinterpret_cast<const uint4*>(x + (b0 + shift) * cin + cc);
            if (b1 >= 0) v1 = *reinterpret_cast<const uint4*>(x + (b1 + shift) * cin + cc);
            const uint32_t a[4] = {v0.x, v0.y, v0.z, v0.w}, c[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                tile_t[cc + 2 * e][pr] = __byte_perm(a[e], c[e], 0x5410);         // channel 2e  : (pixel 2pr, pixel 2pr+1)
                tile_t[cc + 2 * e + 1][pr] = __byte_perm(a[e], c[e], 0x7632);     // channel 2e+1
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_st; i += blockDim.x) {
            const int ci = i / (I2C_PIX / 8), mm = (i % (I2C_PIX / 8)) * 8;
            if (m0 + mm < ldo)
                *reinterpret_cast<uint4*>(out + (static_cast<int64_t>(tap) * cin + ci) * ldo + m0 + mm) =
                    *reinterpret_cast<const uint4*>(&tile_t[ci][mm >> 1]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
im2col_t_stem_kernel(const float* __restrict__ x, bf16* __restrict__ out, int n, int h, int w, int64_t ldo) {
    const int ho = (h - 3) / 2 + 1, wo = (w - 3) / 2 + 1;
    const int64_t m_total = static_cast<int64_t>(n) * ho * wo;
    const int64_t m8 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
    const int k = blockIdx.y;          // 0..31
    if (m8 >= ldo) return;
    bf16 tmp[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float v = 0.f;
        const int64_t m = m8 + e;
        if (k < 27 && m < m_total) {
            const int ci = k / 9, ky = (k - ci * 9) / 3, kx = k % 3;
            const int ox = static_cast<int>(m % wo);
            const int64_t q = m / wo;
            const int oy = static_cast<int>(q % ho);
            const int img = static_cast<int>(q / ho);
            v = x[((static_cast<int64_t>(img) * 3 + ci) * h + 2 * oy + ky) * w + 2 * ox + kx];
        }
        tmp[e] = __float2bfloat16_rn(v);
    }
    *reinterpret_cast<uint4*>(out + static_cast<int64_t>(k) * ldo + m8) = *reinterpret_cast<const uint4*>(tmp);
}

static inline unsigned nblk2(int64_t total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }
static inline int chan_threads(int c) {
    const int c8 = c / 8;
    int lanes = 256 / c8;
    if (lanes < 1) lanes = 1;
    int t = lanes * c8;
    return (t + 31) / 32 * 32;
}

}  // namespace istvt

using namespace istvt;

extern "C" int istvt_bn_stats_fwd(const void* x, float* sum, float* sumsq, int64_t m, int c, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && sum && sumsq && m > 0 && c > 0 && c % 8 == 0 && c <= 2048);
    int64_t blocks = (m + 63) / 64;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    bn_stats_kernel<<<static_cast<unsigned>(blocks), chan_threads(c), 2 * c * sizeof(float),
                      static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x), sum, sumsq, m, c);
    count_launch();
    return launch_status();
}

extern "C" int istvt_bn_finalize_fwd(const float* sum, const float* sumsq, const float* gamma, const float* beta,
                                     float* scale, float* shift, float* mean, float* rstd, float* running_mean,
                                     float* running_var, int64_t m, int c, float eps, float momentum,
                                     istvt_stream_t stream) {
    ISTVT_REQUIRE(sum && sumsq && gamma && beta && scale && shift && mean && rstd && m > 0 && c > 0);
    ISTVT_REQUIRE((running_mean == nullptr) == (running_var == nullptr));
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        sum, sumsq, gamma, beta, scale, shift, mean, rstd, running_mean, running_var, m, c, eps, momentum);
    count_launch();
    return launch_status();
}

extern "C" int istvt_bn_apply_fwd(const void* x, const float* scale, const float* shift, void* y, int64_t m, int c,
                                  int relu, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && scale && shift && y && m > 0 && c > 0 && c % 8 == 0);
    const int64_t n8 = m * (c / 8);
    bn_apply_kernel<<<nblk2(n8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(x), scale, shift, static_cast<bf16*>(y), n8, c / 8, relu);
    count_launch();
    return launch_status();
}

// dgamma / dbeta must hold ONLY this layer's sums when the call returns (zero them once per step): dx uses them.
extern "C" int istvt_bn_bwd(const void* dy, const void* x, const float* scale, const float* shift, const float* mean,
                            const float* rstd, float* dgamma, float* dbeta, void* dx, int64_t m, int c, int relu,
                            istvt_stream_t stream) {
    ISTVT_REQUIRE(dy && x && scale && shift && mean && rstd && dgamma && dbeta && dx && m > 0 && c > 0 && c % 8 == 0 &&
                  c <= 2048);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int64_t blocks = (m + 63) / 64;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    bn_bwd_reduce_kernel<<<static_cast<unsigned>(blocks), chan_threads(c), 2 * c * sizeof(float), st>>>(
        static_cast<const bf16*>(dy), static_cast<const bf16*>(x), scale, shift, mean, rstd, dgamma, dbeta, m, c, relu);
    count_launch();
    // ISTVT_BN_FLAT=1: the flat one-vector-per-thread apply kernel (A/B measurements)
    static const bool flat = []() { const char* e = getenv("ISTVT_BN_FLAT"); return e && atoi(e) != 0; }();
    if (flat) {
        const int64_t n8 = m * (c / 8);
        bn_bwd_apply_kernel<<<nblk2(n8, 256), 256, 0, st>>>(static_cast<const bf16*>(dy), static_cast<const bf16*>(x), scale,
                                                            shift, mean, rstd, dgamma, dbeta, static_cast<bf16*>(dx), n8,
                                                            c / 8, 1.0f / static_cast<float>(m), relu);
    } else {
        const int threads = chan_threads(c);
        const int lanes = (256 / (c / 8)) < 1 ? 1 : 256 / (c / 8);
        int64_t ablocks = (m + 4 * lanes - 1) / (4 * lanes);          // >= 4 rows per thread
        const int64_t acap = static_cast<int64_t>(sm_count()) * 16;
        if (ablocks > acap) ablocks = acap;
        bn_bwd_apply_rows_kernel<<<static_cast<unsigned>(ablocks), threads, 0, st>>>(
            static_cast<const bf16*>(dy), static_cast<const bf16*>(x), scale, shift, mean, rstd, dgamma, dbeta,
            static_cast<bf16*>(dx), m, c, 1.0f / static_cast<float>(m), relu);
    }
    count_launch();
    return launch_status();
}

extern "C" int istvt_pool_add_idx_fwd(const void* x, const void* skip, void* y, const float* pos_emb, float* tokens,
                                      void* argmax, int n, int t_frames, int h, int w, int c, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && skip && argmax && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0);
    ISTVT_REQUIRE((y != nullptr) != (tokens != nullptr));
    ISTVT_REQUIRE(tokens == nullptr || (pos_emb != nullptr && t_frames > 0 && n % t_frames == 0));
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
    const int64_t total = static_cast<int64_t>(n) * ho * wo * (c / 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (tokens == nullptr)
        pool_add_idx_kernel<false><<<nblk2(total, 256), 256, 0, st>>>(
            static_cast<const bf16*>(x), static_cast<const bf16*>(skip), static_cast<bf16*>(y), nullptr, nullptr,
            static_cast<uint8_t*>(argmax), n, h, w, c, ho, wo, 1);
    else
        pool_add_idx_kernel<true><<<nblk2(total, 256), 256, 0, st>>>(
            static_cast<const bf16*>(x), static_cast<const bf16*>(skip), nullptr, pos_emb, tokens,
            static_cast<uint8_t*>(argmax), n, h, w, c, ho, wo, t_frames);
    count_launch();
    return launch_status();
}

extern "C" int istvt_pool_bwd(const void* dy, const void* argmax, void* dx, int n, int h, int w, int c,
                              istvt_stream_t stream) {
    ISTVT_REQUIRE(dy && argmax && dx && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0);
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
    ISTVT_REQUIRE(h <= 65535 && n <= 65535);
    const dim3 grid(static_cast<unsigned>((((w + 1) / 2) * (c / 8) + 255) / 256), static_cast<unsigned>((h + 1) / 2),
                    static_cast<unsigned>(n));
    pool_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(dy), static_cast<const uint8_t*>(argmax), static_cast<bf16*>(dx), n, h, w, c, ho, wo);
    count_launch();
    return launch_status();
}

extern "C" int istvt_token_grad_gather(const float* g, void* d_out, int batch, int t, int tokens_per_frame, int c,
                                       istvt_stream_t stream) {
    ISTVT_REQUIRE(g && d_out && batch > 0 && t > 0 && tokens_per_frame > 1 && c % 8 == 0);
    const int64_t total = static_cast<int64_t>(batch) * t * (tokens_per_frame - 1) * (c / 8);
    token_grad_gather_kernel<<<nblk2(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        g, static_cast<bf16*>(d_out), batch * t, t, tokens_per_frame, c);
    count_launch();
    return launch_status();
}

extern "C" int istvt_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, int n, int h, int w, int c, int relu_in,
                                     istvt_stream_t stream) {
    ISTVT_REQUIRE(x && dy && dw && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0);
    ISTVT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0);
    CUtensorMap tm_x, tm_dy;
    const uint64_t dims[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(w), static_cast<uint64_t>(h),
                              static_cast<uint64_t>(n)};
    const uint64_t strides[3] = {static_cast<uint64_t>(c) * 2, static_cast<uint64_t>(w) * c * 2,
                                 static_cast<uint64_t>(h) * w * c * 2};
    const uint32_t box_x[4] = {DWG_CG, DWG_TW + 2, DWG_TH + 2, 1};
    const uint32_t box_d[4] = {DWG_CG, DWG_TW, DWG_TH, 1};
    int rc = encode_tmap(&tm_x, x, ISTVT_BF16, 4, dims, strides, box_x, 0);
    if (rc != ISTVT_OK) return rc;
    rc = encode_tmap(&tm_dy, dy, ISTVT_BF16, 4, dims, strides, box_d, 0);
    if (rc != ISTVT_OK) return rc;
    ISTVT_CHECK_CUDA(cudaFuncSetAttribute(dwconv_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DWG_SMEM));
    const int64_t total = static_cast<int64_t>(n) * ((c + DWG_CG - 1) / DWG_CG) * ((h + DWG_TH - 1) / DWG_TH) *
                          ((w + DWG_TW - 1) / DWG_TW);
    int64_t grid = static_cast<int64_t>(sm_count()) * 2;
    if (grid > total) grid = total;
    dwconv_wgrad_tma_kernel<<<static_cast<unsigned>(grid), DWG_THREADS, DWG_SMEM, static_cast<cudaStream_t>(stream)>>>(
        tm_x, tm_dy, dw, n, h, w, c, relu_in);
    count_launch();
    return launch_status();
}

extern "C" int istvt_block_input_grad(const void* d_main, const void* x_in, const void* d_skip, void* dx, int n, int h,
                                      int w, int c, int relu_in, istvt_stream_t stream) {
    ISTVT_REQUIRE(d_main && d_skip && dx && (x_in || !relu_in) && n > 0 && h > 0 && w > 0 && c % 8 == 0);
    const int64_t total = static_cast<int64_t>(n) * h * w * (c / 8);
    block_input_grad_kernel<<<nblk2(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(d_main), static_cast<const bf16*>(x_in), static_cast<const bf16*>(d_skip),
        static_cast<bf16*>(dx), n, h, w, c, relu_in);
    count_launch();
    return launch_status();
}

extern "C" int istvt_im2col_t(const void* x, void* out, int n, int h, int w, int cin, int64_t ldo,
                              istvt_stream_t stream) {
    ISTVT_REQUIRE(x && out && n > 0 && h > 2 && w > 2 && cin % 8 == 0 && cin <= 64 && ldo % 8 == 0);
    ISTVT_REQUIRE(ldo >= static_cast<int64_t>(n) * (h - 2) * (w - 2));
    const dim3 grid(static_cast<unsigned>((ldo + I2C_PIX - 1) / I2C_PIX));
    im2col_t_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x),
                                                                         static_cast<bf16*>(out), n, h, w, cin, ldo);
    count_launch();
    return launch_status();
}

extern "C" int istvt_im2col_t_stem(const float* x, void* out, int n, int h, int w, int64_t ldo, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && out && n > 0 && h >= 3 && w >= 3 && ldo % 8 == 0);
    ISTVT_REQUIRE(ldo >= static_cast<int64_t>(n) * ((h - 3) / 2 + 1) * ((w - 3) / 2 + 1));
    const dim3 grid(static_cast<unsigned>((ldo / 8 + 255) / 256), 32);
    im2col_t_stem_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<bf16*>(out), n, h, w, ldo);
    count_launch();
    return launch_status();
}
