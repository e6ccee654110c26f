// HBM-bound kernels of the Xception entry flow, NHWC activations (channel innermost, 16-byte vectors):
//   conv_stem      3x3 s2 conv 3->32 on the NCHW fp32 clip + folded BN + ReLU            (xception.py:194-196)
//   dwconv3x3      depthwise 3x3 p1, ReLU-on-load, TMA halo tiles + rolling accumulators      (xception.py:43,47)
//   subsample2     pixel gather of the stride-2 1x1 skip convolution                     (xception.py:57,94)
//   pool_add       MaxPool2d(3,2,1) + skip add                                           (xception.py:87-88,100)
//   pool_add_tokens  same, writing fp32 tokens + positional embedding                    (+ vivit.py:133-138)
#include "common.cuh"
#include "ptx.cuh"
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
                 T* __restrict__ y, int n, int h, int w, int ho, int wo, int relu) {
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
            for (int e = 0; e < 8; ++e) o[e] = relu ? fmaxf(acc0[c + e], 0.0f) : acc0[c + e];
            store8(yp + c, o);
        }
        if (has2) {
#pragma unroll
            for (int c = 0; c < STEM_CO; c += 8) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = relu ? fmaxf(acc1[c + e], 0.0f) : acc1[c + e];
                store8(yp + STEM_CO + c, o);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// depthwise 3x3 pad 1, HBM-bound.  Persistent CTAs walk (image, channel-group, tile) work items; the
// (TH+2) x (TW+2) x 64-channel halo tile of each item is fetched by ONE TMA load of a 4-D tensor map
// (c, x, y, n) — the pad-1 border and ragged edges are the TMA's out-of-bounds zero fill — into a
// double-buffered smem slot, so the load of tile i+1 overlaps the arithmetic of tile i.
// Thread = (4 channels, one tile column): it walks down the column, reads 3 smem vectors per input
// row and keeps three rolling output-row accumulators in registers (each input row feeds output rows
// r-2, r-1, r with ky = 2, 1, 0), so every smem element is read 3x and every output written once.
// ------------------------------------------------------------------------------------------
constexpr int DW_TW = 16;          // tile width  (outputs)
constexpr int DW_CG = 64;          // channels per work item
constexpr int DW_THREADS = 256;    // 16 channel quads x 16 columns
template <typename T> struct DwCfg;
template <> struct DwCfg<__nv_bfloat16> { static constexpr int TH = 16; };
template <> struct DwCfg<float> { static constexpr int TH = 8; };

__device__ __forceinline__ void lds4(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}
__device__ __forceinline__ void lds4(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

template <typename T>
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv3x3_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ wt, T* __restrict__ y,
                     int n, int h, int w, int c, int relu_in) {
    constexpr int TH = DwCfg<T>::TH;
    constexpr int IN_H = TH + 2, IN_W = DW_TW + 2;
    constexpr uint32_t TILE_BYTES = IN_H * IN_W * DW_CG * sizeof(T);
    extern __shared__ __align__(128) uint8_t dw_smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dw_smem) + 127) & ~uintptr_t(127));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * TILE_BYTES);

    const int tiles_x = (w + DW_TW - 1) / DW_TW;
    const int tiles_y = (h + TH - 1) / TH;
    const int cgroups = (c + DW_CG - 1) / DW_CG;
    const int64_t total = static_cast<int64_t>(n) * cgroups * tiles_y * tiles_x;

    const int tid = threadIdx.x;
    const int cq = tid & 15;          // channel quad inside the 64-channel group
    const int col = tid >> 4;         // tile column

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
        tma_prefetch_desc(&tm_x);
    }
    __syncthreads();

    auto issue = [&](int64_t item, int buf) {
        const int tx = static_cast<int>(item % tiles_x);
        int64_t r = item / tiles_x;
        const int ty = static_cast<int>(r % tiles_y);
        r /= tiles_y;
        const int cg = static_cast<int>(r % cgroups);
        const int img = static_cast<int>(r / cgroups);
        mbar_arrive_expect_tx(&bars[buf], TILE_BYTES);
        tma_load_4d(base + buf * TILE_BYTES, &tm_x, &bars[buf], cg * DW_CG, tx * DW_TW - 1, ty * TH - 1, img);
    };

    int64_t item = blockIdx.x;
    if (tid == 0 && item < total) issue(item, 0);
    int buf = 0;
    uint32_t phase[2] = {0, 0};
    for (; item < total; item += gridDim.x) {
        const int64_t next = item + gridDim.x;
        if (tid == 0 && next < total) issue(next, buf ^ 1);   // slot buf^1 was released by the barrier below

        const int tx = static_cast<int>(item % tiles_x);
        int64_t r = item / tiles_x;
        const int ty = static_cast<int>(r % tiles_y);
        r /= tiles_y;
        const int cg = static_cast<int>(r % cgroups);
        const int img = static_cast<int>(r / cgroups);
        const int ch = cg * DW_CG + cq * 4;
        const bool ch_ok = ch < c;
        const int ox = tx * DW_TW + col;
        const int oy0 = ty * TH;

        float wk[9][4];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float4 t = ch_ok ? __ldg(reinterpret_cast<const float4*>(wt + k * c + ch)) : make_float4(0, 0, 0, 0);
            wk[k][0] = t.x; wk[k][1] = t.y; wk[k][2] = t.z; wk[k][3] = t.w;
        }

        mbar_wait(&bars[buf], phase[buf]);
        phase[buf] ^= 1;

        const T* tile = reinterpret_cast<const T*>(base + buf * TILE_BYTES) + col * DW_CG + cq * 4;
        T* yout = y + ((static_cast<int64_t>(img) * h + oy0) * w + ox) * c + ch;
        const bool st_ok = ch_ok && ox < w;
        float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int rr = 0; rr < IN_H; ++rr) {
            float v[3][4];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                lds4(tile + (rr * IN_W + kx) * DW_CG, v[kx]);
                if (relu_in) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[kx][e] = fmaxf(v[kx][e], 0.0f);
                }
            }
            float a2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a0[e] = fmaf(v[0][e], wk[6][e], a0[e]); a0[e] = fmaf(v[1][e], wk[7][e], a0[e]); a0[e] = fmaf(v[2][e], wk[8][e], a0[e]);
                a1[e] = fmaf(v[0][e], wk[3][e], a1[e]); a1[e] = fmaf(v[1][e], wk[4][e], a1[e]); a1[e] = fmaf(v[2][e], wk[5][e], a1[e]);
                a2[e] = v[0][e] * wk[0][e]; a2[e] = fmaf(v[1][e], wk[1][e], a2[e]); a2[e] = fmaf(v[2][e], wk[2][e], a2[e]);
            }
            if (rr >= 2) {
                const int orow = rr - 2;
                if (st_ok && oy0 + orow < h) store4(yout + static_cast<int64_t>(orow) * w * c, a0);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) { a0[e] = a1[e]; a1[e] = a2[e]; }
        }
        __syncthreads();   // everyone is done reading slot `buf`: it may be refilled next iteration
        buf ^= 1;
    }
}

// (A warp-level tensor-core variant — mma.sync with diagonal B fragments, ldmatrix row pointers into a swizzled
// halo tile, 6 instructions per output instead of ~20 — was built, validated and measured SLOWER: 6.0 vs 4.4 ms
// per step.  Legacy mma.sync on sm_100a peaks near 512 FLOP/clk/SM, 1/16 of tcgen05, and the diagonal trick
// wastes 7/8 of it.  profiles/README.md r1x.)

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

static int conv_stem_launch(const float* x, const float* wt, const float* bias, void* y, int dtype, int n, int h, int w,
                            int cout, int relu, cudaStream_t st) {
    ISTVT_REQUIRE(x && wt && bias && y);
    ISTVT_REQUIRE(cout == STEM_CO && n > 0 && h >= 3 && w >= 3);
    const int ho = (h - 3) / 2 + 1, wo = (w - 3) / 2 + 1;
    const int64_t total = static_cast<int64_t>(n) * ho * ((wo + 1) / 2);
    int64_t blocks = (total + 127) / 128;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 32;
    if (blocks > cap) blocks = cap;
    if (dtype == ISTVT_BF16)
        conv_stem_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 128, 0, st>>>(
            x, wt, bias, static_cast<__nv_bfloat16*>(y), n, h, w, ho, wo, relu);
    else if (dtype == ISTVT_F32)
        conv_stem_kernel<float><<<static_cast<unsigned>(blocks), 128, 0, st>>>(x, wt, bias, static_cast<float*>(y), n,
                                                                                 h, w, ho, wo, relu);
    else
        return ISTVT_ERR_INVALID_ARG;
    count_launch();
    return launch_status();
}

extern "C" int istvt_conv_stem_fwd(const float* x, const float* wt, const float* bias, void* y, int dtype, int n,
                                   int h, int w, int cout, istvt_stream_t stream) {
    return conv_stem_launch(x, wt, bias, y, dtype, n, h, w, cout, 1, static_cast<cudaStream_t>(stream));
}

// Training mode: the raw convolution (+ bias, normally zero) without the ReLU; BatchNorm batch statistics follow.
extern "C" int istvt_conv_stem_raw_fwd(const float* x, const float* wt, const float* bias, void* y, int dtype, int n,
                                       int h, int w, int cout, istvt_stream_t stream) {
    return conv_stem_launch(x, wt, bias, y, dtype, n, h, w, cout, 0, static_cast<cudaStream_t>(stream));
}

template <typename T>
static int launch_dwconv(const void* x, const float* wt, void* y, int n, int h, int w, int c, int relu_in,
                         cudaStream_t st) {
    constexpr int TH = DwCfg<T>::TH;
    constexpr int IN_H = TH + 2, IN_W = DW_TW + 2;
    CUtensorMap tm;
    const uint64_t es = sizeof(T);
    const uint64_t dims[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(w), static_cast<uint64_t>(h),
                              static_cast<uint64_t>(n)};
    const uint64_t strides[3] = {static_cast<uint64_t>(c) * es, static_cast<uint64_t>(w) * c * es,
                                 static_cast<uint64_t>(h) * w * c * es};
    const uint32_t box[4] = {DW_CG, IN_W, IN_H, 1};
    int rc = encode_tmap(&tm, x, sizeof(T) == 2 ? ISTVT_BF16 : ISTVT_F32, 4, dims, strides, box, 0);
    if (rc != ISTVT_OK) return rc;
    const int smem = 2 * IN_H * IN_W * DW_CG * static_cast<int>(sizeof(T)) + 128 + 64;
    auto kern = dwconv3x3_tma_kernel<T>;
    ISTVT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t total = static_cast<int64_t>(n) * ((c + DW_CG - 1) / DW_CG) * ((h + TH - 1) / TH) *
                          ((w + DW_TW - 1) / DW_TW);
    int64_t grid = static_cast<int64_t>(sm_count()) * 2;
    if (grid > total) grid = total;
    kern<<<static_cast<unsigned>(grid), DW_THREADS, smem, st>>>(tm, wt, static_cast<T*>(y), n, h, w, c, relu_in);
    count_launch();
    return launch_status();
}

extern "C" int istvt_dwconv3x3_fwd(const void* x, const float* wt, void* y, int dtype, int n, int h, int w, int c,
                                   int relu_in, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && wt && y);
    ISTVT_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0);
    ISTVT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(wt) & 15) == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == ISTVT_BF16) return launch_dwconv<__nv_bfloat16>(x, wt, y, n, h, w, c, relu_in, st);
    if (dtype == ISTVT_F32) return launch_dwconv<float>(x, wt, y, n, h, w, c, relu_in, st);
    return ISTVT_ERR_INVALID_ARG;
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
