// HBM-bound kernels of the Xception entry flow, NHWC activations (channel innermost, 16-byte vectors):
//   conv_stem      3x3 s2 conv 3->32 on the NCHW fp32 clip + folded BN + ReLU            (xception.py:194-196)
//   dwconv3x3      depthwise 3x3 p1, ReLU-on-load, sliding register window down a column (xception.py:43,47)
//   subsample2     pixel gather of the stride-2 1x1 skip convolution                     (xception.py:57,94)
//   pool_add       MaxPool2d(3,2,1) + skip add                                           (xception.py:87-88,100)
//   pool_add_tokens  same, writing fp32 tokens + positional embedding                    (+ vivit.py:133-138)
#include "common.cuh"
#include "simt_util.cuh"

namespace istvt {

// ------------------------------------------------------------------------------------------
// conv stem: each thread produces 2 horizontally adjacent output pixels x 32 channels.
// Weights [tap = (ci, ky, kx)][co] in shared memory are read by broadcast.
// ------------------------------------------------------------------------------------------
constexpr int STEM_CO = 32;

template <typename T>
__global__ void __launch_bounds__(128)
conv_stem_kernel(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                 T* __restrict__ y, int n, int h, int w, int ho, int wo) {
    __shared__ __align__(16) float sw[27 * STEM_CO];
    __shared__ __align__(16) float sb[STEM_CO];
    // wt is [co][ci][ky][kx] (torch conv weight layout) -> sw[(ci*9 + ky*3 + kx)][co]
    for (int i = threadIdx.x; i < 27 * STEM_CO; i += blockDim.x) {
        const int co = i / 27, tap = i - co * 27;
        sw[tap * STEM_CO + co] = wt[i];
    }
    if (threadIdx.x < STEM_CO) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();

    const int pairs = (wo + 1) >> 1;
    const int64_t total = static_cast<int64_t>(n) * ho * pairs;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int px = static_cast<int>(idx % pairs);
        const int64_t t = idx / pairs;
        const int oy = static_cast<int>(t % ho);
        const int img = static_cast<int>(t / ho);
        const int ox0 = px * 2;
        const bool has2 = ox0 + 1 < wo;

        float acc0[STEM_CO], acc1[STEM_CO];
#pragma unroll
        for (int c = 0; c < STEM_CO; ++c) { acc0[c] = sb[c]; acc1[c] = sb[c]; }

        const float* xin = x + static_cast<int64_t>(img) * 3 * h * w;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float* rowp = xin + (static_cast<int64_t>(ci) * h + (2 * oy + ky)) * w + 2 * ox0;
                float in[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) in[j] = (j < 3 || has2) ? __ldg(rowp + j) : 0.0f;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float* wp = sw + (ci * 9 + ky * 3 + kx) * STEM_CO;
                    const float a0 = in[kx], a1 = in[kx + 2];
#pragma unroll
                    for (int c = 0; c < STEM_CO; c += 4) {
                        const float4 wv = *reinterpret_cast<const float4*>(wp + c);
                        acc0[c] = fmaf(a0, wv.x, acc0[c]); acc0[c + 1] = fmaf(a0, wv.y, acc0[c + 1]);
                        acc0[c + 2] = fmaf(a0, wv.z, acc0[c + 2]); acc0[c + 3] = fmaf(a0, wv.w, acc0[c + 3]);
                        acc1[c] = fmaf(a1, wv.x, acc1[c]); acc1[c + 1] = fmaf(a1, wv.y, acc1[c + 1]);
                        acc1[c + 2] = fmaf(a1, wv.z, acc1[c + 2]); acc1[c + 3] = fmaf(a1, wv.w, acc1[c + 3]);
                    }
                }
            }
        }
        T* yp = y + ((static_cast<int64_t>(img) * ho + oy) * wo + ox0) * STEM_CO;
#pragma unroll
        for (int c = 0; c < STEM_CO; c += 8) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaxf(acc0[c + e], 0.0f);
            store8(yp + c, o);
        }
        if (has2) {
#pragma unroll
            for (int c = 0; c < STEM_CO; c += 8) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = fmaxf(acc1[c + e], 0.0f);
                store8(yp + STEM_CO + c, o);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// depthwise 3x3 pad 1.  Thread = (8 channels, one column x, a strip of DW_ROWS output rows); the 3x3
// window slides down the strip so each new output row costs 3 vector loads.  Threads are ordered
// channel-group fastest then x, so a warp reads contiguous NHWC memory.
// ------------------------------------------------------------------------------------------
constexpr int DW_ROWS = 8;

template <typename T>
__global__ void __launch_bounds__(256)
dwconv3x3_kernel(const T* __restrict__ x, const float* __restrict__ wt, T* __restrict__ y, int n, int h, int w,
                 int c, int relu_in) {
    const int c8 = c >> 3;
    const int strips = (h + DW_ROWS - 1) / DW_ROWS;
    const int64_t total = static_cast<int64_t>(n) * strips * w * c8;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cg = static_cast<int>(idx % c8);
    int64_t t = idx / c8;
    const int xo = static_cast<int>(t % w);
    t /= w;
    const int strip = static_cast<int>(t % strips);
    const int img = static_cast<int>(t / strips);
    const int ch = cg * 8;
    const int y0 = strip * DW_ROWS;

    float wk[9][8];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(wt + k * c + ch));
        const float4 b = __ldg(reinterpret_cast<const float4*>(wt + k * c + ch + 4));
        wk[k][0] = a.x; wk[k][1] = a.y; wk[k][2] = a.z; wk[k][3] = a.w;
        wk[k][4] = b.x; wk[k][5] = b.y; wk[k][6] = b.z; wk[k][7] = b.w;
    }

    const T* xin = x + static_cast<int64_t>(img) * h * w * c + ch;
    T* yout = y + static_cast<int64_t>(img) * h * w * c + ch;
    const bool left = xo > 0, right = xo + 1 < w;

    // rows[r][dx][8]: r = 0 -> row y-1, 1 -> row y, 2 -> row y+1
    float win[3][3][8];
    auto load_row = [&](int yy, float (&dst)[3][8]) {
        const bool row_ok = (yy >= 0) && (yy < h);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const bool ok = row_ok && (dx == 1 || (dx == 0 ? left : right));
            if (ok) {
                load8(xin + (static_cast<int64_t>(yy) * w + (xo + dx - 1)) * c, dst[dx]);
                if (relu_in) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) dst[dx][e] = fmaxf(dst[dx][e], 0.0f);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) dst[dx][e] = 0.0f;
            }
        }
    };
    load_row(y0 - 1, win[0]);
    load_row(y0, win[1]);
#pragma unroll
    for (int r = 0; r < DW_ROWS; ++r) {
        const int yy = y0 + r;
        if (yy >= h) break;
        load_row(yy + 1, win[2]);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(win[ky][kx][e], wk[ky * 3 + kx][e], acc[e]);
        store8(yout + (static_cast<int64_t>(yy) * w + xo) * c, acc);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                win[0][dx][e] = win[1][dx][e];
                win[1][dx][e] = win[2][dx][e];
            }
    }
}

// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
subsample2_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w, int c, int ho, int wo) {
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
    float v[8];
    load8(x + ((static_cast<int64_t>(img) * h + 2 * oy) * w + 2 * ox) * c + cg * 8, v);
    store8(y + idx * 8, v);
}

// ------------------------------------------------------------------------------------------
// maxpool 3x3 s2 p1 (+ skip).  TOKENS = true writes fp32 tokens[b, f+1, 1+p, :] + pos_emb[f, 1+p, :].
// ------------------------------------------------------------------------------------------
template <typename T, bool TOKENS>
__global__ void __launch_bounds__(256)
pool_add_kernel(const T* __restrict__ x, const T* __restrict__ skip, T* __restrict__ y,
                const float* __restrict__ pos_emb, float* __restrict__ tokens, int n, int h, int w, int c, int ho,
                int wo, int t_frames) {
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
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
    const T* xin = x + static_cast<int64_t>(img) * h * w * c + ch;
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
            for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
        }
    }
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

static inline unsigned blocks_for(int64_t total, int threads) {
    return static_cast<unsigned>((total + threads - 1) / threads);
}

}  // namespace istvt

using namespace istvt;

extern "C" int istvt_conv_stem_fwd(const float* x, const float* wt, const float* bias, void* y, int dtype, int n,
                                   int h, int w, int cout, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && wt && bias && y);
    ISTVT_REQUIRE(cout == STEM_CO && n > 0 && h >= 3 && w >= 3);
    const int ho = (h - 3) / 2 + 1, wo = (w - 3) / 2 + 1;
    const int64_t total = static_cast<int64_t>(n) * ho * ((wo + 1) / 2);
    int64_t blocks = (total + 127) / 128;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 32;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == ISTVT_BF16)
        conv_stem_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 128, 0, st>>>(
            x, wt, bias, static_cast<__nv_bfloat16*>(y), n, h, w, ho, wo);
    else if (dtype == ISTVT_F32)
        conv_stem_kernel<float><<<static_cast<unsigned>(blocks), 128, 0, st>>>(x, wt, bias, static_cast<float*>(y), n,
                                                                                 h, w, ho, wo);
    else
        return ISTVT_ERR_INVALID_ARG;
    count_launch();
    return launch_status();
}

extern "C" int istvt_dwconv3x3_fwd(const void* x, const float* wt, void* y, int dtype, int n, int h, int w, int c,
                                   int relu_in, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && wt && y);
    ISTVT_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0);
    const int strips = (h + DW_ROWS - 1) / DW_ROWS;
    const int64_t total = static_cast<int64_t>(n) * strips * w * (c / 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == ISTVT_BF16)
        dwconv3x3_kernel<__nv_bfloat16><<<blocks_for(total, 256), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), wt, static_cast<__nv_bfloat16*>(y), n, h, w, c, relu_in);
    else if (dtype == ISTVT_F32)
        dwconv3x3_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>(static_cast<const float*>(x), wt,
                                                                         static_cast<float*>(y), n, h, w, c, relu_in);
    else
        return ISTVT_ERR_INVALID_ARG;
    count_launch();
    return launch_status();
}

extern "C" int istvt_subsample2_fwd(const void* x, void* y, int dtype, int n, int h, int w, int c,
                                    istvt_stream_t stream) {
    ISTVT_REQUIRE(x && y);
    ISTVT_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0);
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
    const int64_t total = static_cast<int64_t>(n) * ho * wo * (c / 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == ISTVT_BF16)
        subsample2_kernel<__nv_bfloat16><<<blocks_for(total, 256), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n, h, w, c, ho, wo);
    else if (dtype == ISTVT_F32)
        subsample2_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>(static_cast<const float*>(x),
                                                                          static_cast<float*>(y), n, h, w, c, ho, wo);
    else
        return ISTVT_ERR_INVALID_ARG;
    count_launch();
    return launch_status();
}

extern "C" int istvt_pool_add_fwd(const void* x, const void* skip, void* y, int dtype, int n, int h, int w, int c,
                                  istvt_stream_t stream) {
    ISTVT_REQUIRE(x && skip && y);
    ISTVT_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0);
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
    const int64_t total = static_cast<int64_t>(n) * ho * wo * (c / 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == ISTVT_BF16)
        pool_add_kernel<__nv_bfloat16, false><<<blocks_for(total, 256), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(skip),
            static_cast<__nv_bfloat16*>(y), nullptr, nullptr, n, h, w, c, ho, wo, 1);
    else if (dtype == ISTVT_F32)
        pool_add_kernel<float, false><<<blocks_for(total, 256), 256, 0, st>>>(
            static_cast<const float*>(x), static_cast<const float*>(skip), static_cast<float*>(y), nullptr, nullptr,
            n, h, w, c, ho, wo, 1);
    else
        return ISTVT_ERR_INVALID_ARG;
    count_launch();
    return launch_status();
}

extern "C" int istvt_pool_add_tokens_fwd(const void* x, const void* skip, const float* pos_emb, float* tokens,
                                         int dtype, int batch, int t, int h, int w, int c, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && skip && pos_emb && tokens);
    ISTVT_REQUIRE(batch > 0 && t > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0);
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
    const int n = batch * t;
    const int64_t total = static_cast<int64_t>(n) * ho * wo * (c / 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == ISTVT_BF16)
        pool_add_kernel<__nv_bfloat16, true><<<blocks_for(total, 256), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(skip), nullptr, pos_emb, tokens,
            n, h, w, c, ho, wo, t);
    else if (dtype == ISTVT_F32)
        pool_add_kernel<float, true><<<blocks_for(total, 256), 256, 0, st>>>(
            static_cast<const float*>(x), static_cast<const float*>(skip), nullptr, pos_emb, tokens, n, h, w, c, ho,
            wo, t);
    else
        return ISTVT_ERR_INVALID_ARG;
    count_launch();
    return launch_status();
}
