// HBM-bound kernels of the Xception entry flow, NHWC activations (channel innermost, 16-byte vectors):
//   conv_stem      3x3 s2 conv 3->32 on the NCHW fp32 clip + folded BN + ReLU            (xception.py:194-196)
//   dwconv3x3      depthwise 3x3 p1, ReLU-on-load, TMA halo tiles + rolling accumulators      (xception.py:43,47)
//   subsample2     pixel gather of the stride-2 1x1 skip convolution                     (xception.py:57,94)
//   pool_add       MaxPool2d(3,2,1) + skip add                                           (xception.py:87-88,100)
//   pool_add_tokens  same, writing fp32 tokens + positional embedding                    (+ vivit.py:133-138)
#include "common.cuh"
#include "ptx.cuh"
#include "simt_util.cuh"

#include <stdlib.h>

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
// conv stem on decoded frames: uint8 NHWC [n, h, w, 3] in, same arithmetic.  The input normalisation
// ((x / 255 - mean_c) / std_c) is affine per input channel and conv1 has no padding, so the host folds it into
// `wt` / `bias` exactly like the BatchNorm scale; the kernel only converts bytes.  Each thread reads the 3 x 15
// contiguous bytes under its two output pixels (7 aligned 16-bit loads + 1 byte per row).  Cuts the H2D copy and
// the stem's HBM read 4x against the fp32 NCHW clip.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
conv_stem_u8_kernel(const uint8_t* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                    T* __restrict__ y, int n, int h, int w, int ho, int wo) {
    __shared__ __align__(16) float sw[27 * STEM_CO];
    __shared__ __align__(16) float sb[STEM_CO];
    // wt is [co][ci][ky][kx] -> sw[(ky*9 + px*3 + ci)][co] for px = kx: byte order of an input row is (pixel, channel)
    for (int i = threadIdx.x; i < 27 * STEM_CO; i += blockDim.x) {
        const int co = i / 27, r = i - co * 27;
        const int ci = r / 9, ky = (r - ci * 9) / 3, kx = r % 3;
        sw[(ky * 9 + kx * 3 + ci) * STEM_CO + co] = wt[i];
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

#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            // 5 pixels x 3 channels = 15 bytes at an even offset (6 * ox0 + 3 * w * row): 16-bit loads are aligned
            // when the row pitch 3 * w is even; otherwise fall back to byte loads for that row
            const uint8_t* rowp = x + ((static_cast<int64_t>(img) * h + (2 * oy + ky)) * w + 2 * ox0) * 3;
            float in[15];
            const int nb = has2 ? 15 : 9;          // bytes under this thread's outputs; never read past them
            if ((reinterpret_cast<uintptr_t>(rowp) & 1) == 0) {
                const int n16 = nb >> 1;           // 7 or 4 aligned 16-bit loads, then the odd last byte
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const uint32_t v = j < n16 ? __ldg(reinterpret_cast<const unsigned short*>(rowp) + j) : 0u;
                    in[2 * j] = static_cast<float>(v & 0xffu);
                    in[2 * j + 1] = static_cast<float>(v >> 8);
                }
                const float last = static_cast<float>(__ldg(rowp + nb - 1));
                if (has2) in[14] = last; else { in[8] = last; in[14] = 0.0f; }
            } else {
#pragma unroll
                for (int j = 0; j < 15; ++j) in[j] = j < nb ? static_cast<float>(__ldg(rowp + j)) : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) {     // q = kx * 3 + ci
                const float* wp = sw + (ky * 9 + q) * STEM_CO;
                const float a0 = in[q], a1 = in[q + 6];
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

// ------------------------------------------------------------------------------------------
// depthwise 3x3 pad 1, bf16 production kernel: column strips walked top to bottom.
//
// ncu of the tile kernel above (profiles/README.md r1k) showed it instruction-issue bound, not HBM bound: 64 % of
// the issue slots at 2.7-3.1 TB/s, ~34 thread-instructions per output once the (TH+2)/TH halo rows, the 16x16
// tile padding of the 147 / 74 / 37-wide maps (up to 1.68x) and the 3x re-load + ReLU + unpack of every input
// element are counted.  This kernel cuts the count to ~14 per output:
//   * work item = (image, 64-channel group, column strip, row segment); a dedicated producer warp streams the
//     strip through a ring of DS_R-row TMA slabs (4-D map, zero-filled pad-1 border and ragged edges), so there
//     is no per-tile restart and the only halo rows are the 2 at each segment boundary;
//   * thread = 4 channels x TWO adjacent columns: 4 smem vectors per input row feed 2 outputs (was 3 per 1),
//     ReLU on packed bf16 pairs (HMNMX2), three rolling accumulator sets renamed by the 6-row unroll (no moves);
//   * strips are sized to the map (38 outputs: 147 -> 4, 74 -> 2, 37 -> 1 strips, <= 3.4 % padding) and row
//     segments are chosen so that the persistent grid sees >= 12 waves of items.
// HBM floor: (2 + 2) B per output; the FMA pipe needs 9 of the ~14 issue slots.
// ------------------------------------------------------------------------------------------
constexpr int DS_CG = 64;            // channels per item (one 128-byte line per pixel)
// input rows per ring stage (multiple of 3: the accumulator roles realign) x ring depth: 6 x 3 or 3 x 6 (same bytes)
constexpr int DS_MAX_PAIRS = 19;     // column pairs per strip -> strips of <= 38 outputs
constexpr int DS_MAX_THREADS = ((16 * DS_MAX_PAIRS + 31) / 32) * 32;   // 320 (a producer warp would cap the
                                                                        // kernel at 80 registers and spill)

// predicated 8-byte store of 4 fp32 values rounded to bf16 (no branch, no divergence bookkeeping in the hot loop)
__device__ __forceinline__ void stg_bf16x4_pred(__nv_bfloat16* p, const float (&v)[4], bool pred) {
    const uint32_t lo = pack_bf16x2(v[0], v[1]), hi = pack_bf16x2(v[2], v[3]);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.global.v2.b32 [%0], {%1, %2};\n\t}" ::"l"(p), "r"(lo),
                 "r"(hi), "r"(static_cast<uint32_t>(pred))
                 : "memory");
}

struct DwStripPlan {
    int pairs, strips, segs, seg_h, cgroups, warps;
    int64_t items;
};

template <bool RELU, int DS_R, int DS_STAGES>
__global__ void __launch_bounds__(DS_MAX_THREADS, 2)
dwconv3x3_strip_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ wt,
                       __nv_bfloat16* __restrict__ y, int n, int h, int w, int c, const DwStripPlan plan) {
    extern __shared__ __align__(128) uint8_t dw_smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dw_smem) + 127) & ~uintptr_t(127));
    const int in_w = 2 * plan.pairs + 2;
    const uint32_t stage_bytes = static_cast<uint32_t>(DS_R * in_w * DS_CG * 2);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + DS_STAGES * stage_bytes);
    uint64_t* empty_bar = full_bar + DS_STAGES;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int n_st = (plan.seg_h + 2 + DS_R - 1) / DS_R;          // ring stages per item
    const int my_items = plan.items > static_cast<int64_t>(blockIdx.x)
                             ? static_cast<int>((plan.items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const int total_g = my_items * n_st;                           // ring stages this CTA consumes

    if (tid == 0) {
        for (int s = 0; s < DS_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], plan.warps);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tm_x);
    }
    __syncthreads();

    // item -> (img, cg, strip, seg); seg fastest so that a CTA's consecutive items continue down the same strip
    auto decode = [&](int64_t item, int& img, int& cg, int& x0, int& y0) {
        const int seg = static_cast<int>(item % plan.segs);
        int64_t r = item / plan.segs;
        const int strip = static_cast<int>(r % plan.strips);
        r /= plan.strips;
        cg = static_cast<int>(r % plan.cgroups);
        img = static_cast<int>(r / plan.cgroups);
        x0 = strip * 2 * plan.pairs;
        y0 = seg * plan.seg_h;
    };
    // thread 0 only: TMA load of ring stage g (slot g % DS_STAGES), after every warp released the slot's previous use
    auto produce = [&](int g) {
        if (g >= total_g) return;
        const int k = g / n_st, st = g - k * n_st;
        int img, cg, x0, y0;
        decode(static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(k) * gridDim.x, img, cg, x0, y0);
        const int slot = g % DS_STAGES;
        const int use = g / DS_STAGES;
        if (use > 0) mbar_wait(&empty_bar[slot], (use - 1) & 1);
        mbar_arrive_expect_tx(&full_bar[slot], stage_bytes);
        tma_load_4d(base + slot * stage_bytes, &tm_x, &full_bar[slot], cg * DS_CG, x0 - 1, y0 - 1 + st * DS_R, img);
    };
    if (tid == 0) {
        for (int g = 0; g < DS_STAGES - 1; ++g) produce(g);
    }

    const int cq = tid & 15;                       // channel quad inside the group
    const int pair = tid >> 4;                     // column pair inside the strip
    const bool active = pair < plan.pairs;         // the last warp may be half empty
    const int pair_c = active ? pair : 0;
    int stage = 0;
    uint32_t phase = 0;
    int g = 0;
    for (int k = 0; k < my_items; ++k) {
        int img, cg, x0, y0;
        decode(static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(k) * gridDim.x, img, cg, x0, y0);
        const int ch = cg * DS_CG + cq * 4;
        const bool ch_ok = active && ch < c;
        const int ox = x0 + 2 * pair_c;
        const bool st0 = ch_ok && ox < w, st1 = ch_ok && ox + 1 < w;
        const int y_end = min(y0 + plan.seg_h, h);

        // fp32 PAIRS: the multiply-adds are FFMA2 (fma.rn.f32x2, sm_100), two per issue slot — the kernel was issue bound
        // (62 % issue utilisation with 9 of ~14 slots per output on the FMA stream, profiles/README.md r3e)
        f32x2_t wk[9][2];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const float4 t = ch_ok ? __ldg(reinterpret_cast<const float4*>(wt + q * c + ch)) : make_float4(0, 0, 0, 0);
            wk[q][0] = f32x2_make(t.x, t.y);
            wk[q][1] = f32x2_make(t.z, t.w);
        }
        // output pointer of "input row r - 2"; advanced by one image row per input row (no per-store multiplies)
        const int64_t pitch = static_cast<int64_t>(w) * c;
        __nv_bfloat16* yp = y + (static_cast<int64_t>(img) * h * w + ox) * c + ch + static_cast<int64_t>(y0 - 2) * pitch;
        const uint32_t rows_valid = static_cast<uint32_t>(y_end - y0);
        // three accumulator sets; at input row r, set r % 3 is fresh (ky = 0), (r-1) % 3 takes ky = 1 and (r-2) % 3
        // takes ky = 2 and is complete (output row r - 2).  Rows 0 / 1 of an item leave garbage in the sets that
        // would belong to output rows -2 / -1: never stored.
        f32x2_t acc[3][2][2];
        for (int st = 0; st < n_st; ++st, ++g) {
            if (tid == 0) produce(g + DS_STAGES - 1);
            __syncwarp();
            mbar_wait(&full_bar[stage], phase);
            const uint32_t srow = smem_u32(base) + stage * stage_bytes + (2 * pair_c) * (DS_CG * 2) + cq * 8;
            const uint32_t rbase = static_cast<uint32_t>(st * DS_R - 2);
#pragma unroll
            for (int i = 0; i < DS_R; ++i) {
                if (st * DS_R + i >= static_cast<int>(rows_valid) + 2) break;   // slab rows past the segment's last input row (uniform)
                f32x2_t f[4][2];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint2 t;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(t.x), "=r"(t.y) : "r"(srow + (i * in_w + q) * (DS_CG * 2)));
                    if (RELU) {
                        const __nv_bfloat162 z = __float2bfloat162_rn(0.0f);
                        __nv_bfloat162 lo = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&t.x), z);
                        __nv_bfloat162 hi = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&t.y), z);
                        t.x = *reinterpret_cast<uint32_t*>(&lo);
                        t.y = *reinterpret_cast<uint32_t*>(&hi);
                    }
                    f[q][0] = f32x2_make_bits(t.x << 16, t.x & 0xffff0000u);
                    f[q][1] = f32x2_make_bits(t.y << 16, t.y & 0xffff0000u);
                }
                const int s_new = i % 3, s_mid = (i + 2) % 3, s_old = (i + 1) % 3;   // rows r, r-1, r-2 (DS_R % 3 == 0)
#pragma unroll
                for (int o = 0; o < 2; ++o) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        f32x2_t a = acc[s_old][o][e];
                        a = f32x2_fma(f[o][e], wk[6][e], a); a = f32x2_fma(f[o + 1][e], wk[7][e], a); a = f32x2_fma(f[o + 2][e], wk[8][e], a);
                        acc[s_old][o][e] = a;
                        f32x2_t b = acc[s_mid][o][e];
                        b = f32x2_fma(f[o][e], wk[3][e], b); b = f32x2_fma(f[o + 1][e], wk[4][e], b); b = f32x2_fma(f[o + 2][e], wk[5][e], b);
                        acc[s_mid][o][e] = b;
                        f32x2_t d = f32x2_fma(f[o][e], wk[0][e], 0ull);
                        d = f32x2_fma(f[o + 1][e], wk[1][e], d); d = f32x2_fma(f[o + 2][e], wk[2][e], d);
                        acc[s_new][o][e] = d;
                    }
                }
                // input row r = st * DS_R + i of the segment completes output row y0 + r - 2
                const bool row_ok = rbase + i < rows_valid;        // unsigned: also false for r < 2
                float o0[4], o1[4];
                f32x2_split(acc[s_old][0][0], o0[0], o0[1]); f32x2_split(acc[s_old][0][1], o0[2], o0[3]);
                f32x2_split(acc[s_old][1][0], o1[0], o1[1]); f32x2_split(acc[s_old][1][1], o1[2], o1[3]);
                stg_bf16x4_pred(yp, o0, row_ok && st0);
                stg_bf16x4_pred(yp + c, o1, row_ok && st1);
                yp += pitch;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == DS_STAGES) { stage = 0; phase ^= 1; }
        }
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
    const T* xin = x + static_cast<int64_t>(img) * h * w * c + ch;
    if constexpr (sizeof(T) == 2) {
        // bf16: the maximum is exact in bf16, so the 9 taps are reduced on PACKED pairs (4 HMNMX2 per 16-byte load
        // instead of 8 unpacks + 8 FMNMX: the fp32 version was issue-bound at 63 % of the HBM peak, profiles r1k/r1w)
        const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
        __nv_bfloat162 pm[4] = {ninf, ninf, ninf, ninf};
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = 2 * oy - 1 + dy;
            if (iy < 0 || iy >= h) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = 2 * ox - 1 + dx;
                if (ix < 0 || ix >= w) continue;
                const uint4 t = __ldg(reinterpret_cast<const uint4*>(xin + (static_cast<int64_t>(iy) * w + ix) * c));
                pm[0] = __hmax2(pm[0], *reinterpret_cast<const __nv_bfloat162*>(&t.x));
                pm[1] = __hmax2(pm[1], *reinterpret_cast<const __nv_bfloat162*>(&t.y));
                pm[2] = __hmax2(pm[2], *reinterpret_cast<const __nv_bfloat162*>(&t.z));
                pm[3] = __hmax2(pm[3], *reinterpret_cast<const __nv_bfloat162*>(&t.w));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(pm[i]);
            m[2 * i] = f.x;
            m[2 * i + 1] = f.y;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
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

// conv_stem_tc.cu: the bf16-output stem as a TF32 implicit GEMM on tcgen05
int conv_stem_tc_launch(const void* x, bool u8, const float* wt, const float* bias, void* y, int n, int h, int w,
                        int relu, cudaStream_t st);

}  // namespace istvt

using namespace istvt;

// ISTVT_STEM_TC=0: the SIMT stem kernels also for bf16 output (A/B measurements)
static bool stem_tc_enabled() {
    static const bool on = []() { const char* e = getenv("ISTVT_STEM_TC"); return !e || atoi(e) != 0; }();
    return on;
}

static int conv_stem_launch(const float* x, const float* wt, const float* bias, void* y, int dtype, int n, int h, int w,
                            int cout, int relu, cudaStream_t st) {
    ISTVT_REQUIRE(x && wt && bias && y);
    ISTVT_REQUIRE(cout == STEM_CO && n > 0 && h >= 3 && w >= 3);
    const int ho = (h - 3) / 2 + 1, wo = (w - 3) / 2 + 1;
    const int64_t total = static_cast<int64_t>(n) * ho * ((wo + 1) / 2);
    int64_t blocks = (total + 127) / 128;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 32;
    if (blocks > cap) blocks = cap;
    if (dtype == ISTVT_BF16 && stem_tc_enabled()) return conv_stem_tc_launch(x, false, wt, bias, y, n, h, w, relu, st);
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

extern "C" int istvt_conv_stem_u8_fwd(const uint8_t* x, const float* wt, const float* bias, void* y, int dtype, int n,
                                      int h, int w, int cout, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && wt && bias && y);
    ISTVT_REQUIRE(cout == STEM_CO && n > 0 && h >= 3 && w >= 3);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ho = (h - 3) / 2 + 1, wo = (w - 3) / 2 + 1;
    const int64_t total = static_cast<int64_t>(n) * ho * ((wo + 1) / 2);
    int64_t blocks = (total + 127) / 128;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 32;
    if (blocks > cap) blocks = cap;
    if (dtype == ISTVT_BF16 && stem_tc_enabled()) return conv_stem_tc_launch(x, true, wt, bias, y, n, h, w, 1, st);
    if (dtype == ISTVT_BF16)
        conv_stem_u8_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 128, 0, st>>>(
            x, wt, bias, static_cast<__nv_bfloat16*>(y), n, h, w, ho, wo);
    else if (dtype == ISTVT_F32)
        conv_stem_u8_kernel<float><<<static_cast<unsigned>(blocks), 128, 0, st>>>(x, wt, bias, static_cast<float*>(y), n,
                                                                                    h, w, ho, wo);
    else
        return ISTVT_ERR_INVALID_ARG;
    count_launch();
    return launch_status();
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

// strip / segment plan of the bf16 kernel for an [n, h, w, c] map on a persistent grid of `grid` CTAs
static DwStripPlan plan_dw_strips(int n, int h, int w, int c, int64_t grid) {
    DwStripPlan pl{};
    pl.strips = (w + 2 * DS_MAX_PAIRS - 1) / (2 * DS_MAX_PAIRS);
    const int tw = (w + pl.strips - 1) / pl.strips;
    pl.pairs = (tw + 1) / 2;
    pl.cgroups = (c + DS_CG - 1) / DS_CG;
    const int64_t base_items = static_cast<int64_t>(n) * pl.cgroups * pl.strips;
    int64_t segs = (12 * grid + base_items - 1) / base_items;      // >= 12 waves of items ...
    const int64_t max_segs = (h + 17) / 18;                         // ... but segments of >= 18 rows (2 halo rows each)
    if (segs > max_segs) segs = max_segs;
    if (segs < 1) segs = 1;
    pl.seg_h = static_cast<int>((h + segs - 1) / segs);
    pl.segs = (h + pl.seg_h - 1) / pl.seg_h;
    pl.warps = (16 * pl.pairs + 31) / 32;
    pl.items = base_items * pl.segs;
    return pl;
}

static int launch_dwconv_strips(const void* x, const float* wt, void* y, int n, int h, int w, int c, int relu_in,
                                cudaStream_t st) {
    const int64_t grid_max = static_cast<int64_t>(sm_count()) * 2;
    const DwStripPlan pl = plan_dw_strips(n, h, w, c, grid_max);
    const int in_w = 2 * pl.pairs + 2;
    CUtensorMap tm;
    const uint64_t dims[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(w), static_cast<uint64_t>(h),
                              static_cast<uint64_t>(n)};
    const uint64_t strides[3] = {static_cast<uint64_t>(c) * 2, static_cast<uint64_t>(w) * c * 2,
                                 static_cast<uint64_t>(h) * w * c * 2};
    // (3-row stages x 6 instead of 6-row stages x 3 — finer hand-off, same smem — measured slower: 3.30 vs 2.77 ms
    //  over the six C2 layers, profiles/README.md r3f.)
    constexpr int rows = 6, stages = 3;
    const uint32_t box[4] = {DS_CG, static_cast<uint32_t>(in_w), static_cast<uint32_t>(rows), 1};
    int rc = encode_tmap(&tm, x, ISTVT_BF16, 4, dims, strides, box, 0);
    if (rc != ISTVT_OK) return rc;
    const int smem = stages * rows * in_w * DS_CG * 2 + 128 + 64;
    const int threads = pl.warps * 32;
    const int64_t grid = pl.items < grid_max ? pl.items : grid_max;
    if (relu_in) {
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_strip_kernel<true, rows, stages>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        dwconv3x3_strip_kernel<true, rows, stages><<<static_cast<unsigned>(grid), threads, smem, st>>>(
            tm, wt, static_cast<__nv_bfloat16*>(y), n, h, w, c, pl);
    } else {
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_strip_kernel<false, rows, stages>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        dwconv3x3_strip_kernel<false, rows, stages><<<static_cast<unsigned>(grid), threads, smem, st>>>(
            tm, wt, static_cast<__nv_bfloat16*>(y), n, h, w, c, pl);
    }
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
    if (dtype == ISTVT_BF16) {
        // ISTVT_DW_TILES=1 selects the older 16x16-tile kernel (A/B measurements only)
        static const bool tiles = []() { const char* e = getenv("ISTVT_DW_TILES"); return e && atoi(e) != 0; }();
        if (tiles) return launch_dwconv<__nv_bfloat16>(x, wt, y, n, h, w, c, relu_in, st);
        return launch_dwconv_strips(x, wt, y, n, h, w, c, relu_in, st);
    }
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
