// SeparableConv2d (network/xception.py:39-49: depthwise 3x3 pad 1, then pointwise 1x1) + the folded BatchNorm that
// follows it (xception.py:69-75) + optional ReLU, in ONE kernel: the depthwise result never goes to HBM.
//
// Unfused (istvt_dwconv3x3_fwd + istvt_gemm_fwd) the depthwise output of the entry flow's first two blocks is written
// and read back once: 12.5 MB per frame, 9.6 GB per C2 step ("keeping the depthwise output on-chip
// removes ~30 MB/frame").  Here:
//   * work item = (image, column strip of <= 38 pixels, block of 3 output rows) = <= 114 pixels = the TMEM lanes of
//     one accumulator tile [128 x N]; a persistent CTA walks items down a strip;
//   * per 64-channel group: a TMA box (5 rows x (W + 2) columns x 64 channels, zero-filled pad-1 border) lands in a
//     ring slot; 10 SIMT warps (thread = 4 channels x 2 adjacent columns x 3 rows, fp32 accumulation, optional ReLU on
//     the packed input) compute the depthwise outputs and write them as bf16 rows of a K-major SWIZZLE_128B A tile
//     [pixel][64 channels] — exactly what the dwconv kernel would have stored;
//   * one thread issues tcgen05.mma (M 128, N = Cout <= 256, K 64 per group) against the pointwise weights, which stay
//     in shared memory for the life of the CTA; the A tile is double buffered, so the MMAs of group g overlap the
//     depthwise arithmetic of group g + 1;
//   * the bias is added by the tensor core: every accumulator starts as ones[128 x 16] . B_bias[N x 16]^T, where the
//     first two k columns of B_bias hold the bf16 head and tail of the fp32 bias (hi + lo: 16 mantissa bits, far below
//     the bf16 rounding of the output) — no bias loads or adds in the epilogue;
//   * 4 epilogue warps (lane = pixel) read the accumulator (double buffered in TMEM: the epilogue of item i overlaps
//     item i + 1), apply the ReLU, and stage 64 output channels at a time in one of two swizzled slabs that ONE 4-D
//     TMA store (channels, x, y, image) writes back — image borders and ragged strips are clipped by the tensor map.
//     (First version: bias through 32 loads + 64 adds per thread and group, one slab, two barriers per group — the
//     epilogue, not the depthwise arithmetic, set the pace: profiles/r6x_sepconv_fused_ncu_source.txt.)
// HBM traffic = the input read once (+ 2 halo rows of 5, served by L2) and the output written once.
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace istvt {

constexpr int SF_CG = 64;                 // channels per K group (one 128-byte line per pixel)
constexpr int SF_ROWS = 3;                // output rows per item
constexpr int SF_IN_ROWS = SF_ROWS + 2;
constexpr int SF_MAX_PAIRS = 19;          // column pairs per strip: <= 38 columns x 3 rows = 114 lanes
constexpr int SF_DW_WARPS = 10;           // 16 channel quads x 19 pairs = 304 threads
constexpr int SF_EPI_WARPS = 4;
constexpr int SF_WARP_PROD = SF_DW_WARPS, SF_WARP_MMA = SF_DW_WARPS + 1, SF_WARP_EPI0 = SF_DW_WARPS + 2;
constexpr int SF_THREADS = 32 * (SF_DW_WARPS + 2 + SF_EPI_WARPS);      // 512
constexpr int SF_A_BYTES = 128 * SF_CG * 2;                            // 16 KB: 128 pixels x 64 bf16
constexpr int SF_SLAB_BYTES = 128 * 128;                               // 16 KB: 128 pixels x 64 bf16 output channels
constexpr int SF_ONES_BYTES = 128 * 32;                                // 4 KB: A tile of the bias MMA, [128 x 16] SW32
constexpr int SF_MAX_STAGES = 4;

struct SepFusedPlan {
    int pairs, strips, rblocks, cgroups, stages;
    int64_t items;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

template <bool RELU_IN>
__global__ void __launch_bounds__(SF_THREADS, 1)
sepconv_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                     const __grid_constant__ CUtensorMap tm_y, const float* __restrict__ dw, const float* __restrict__ bias,
                     int h, int w, int c, int n_out, int act, const SepFusedPlan plan) {
    extern __shared__ uint8_t sf_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sf_raw) + 1023) & ~uintptr_t(1023));
    const int in_w = 2 * plan.pairs + 2;
    const uint32_t stage_bytes = static_cast<uint32_t>(SF_IN_ROWS * in_w * SF_CG * 2);
    const uint32_t w_chunk = static_cast<uint32_t>(n_out) * 128u;                    // one K group of the weights
    uint8_t* s_a = smem;                                          // 2 x 16 KB
    uint8_t* s_slab = s_a + 2 * SF_A_BYTES;                       // 2 x 16 KB
    uint8_t* s_ones = s_slab + 2 * SF_SLAB_BYTES;                 // 4 KB: [128 x 16] K-major SW32, k columns 0 and 1 = 1.0
    uint8_t* s_bias = s_ones + SF_ONES_BYTES;                     // n_out x 32 B: [n_out x 16] SW32, k 0 / 1 = bias hi / lo
    uint8_t* s_w = s_bias + 256 * 32;                             // cgroups x n_out x 128 B
    uint8_t* s_in = s_w + plan.cgroups * w_chunk;                 // stages x stage_bytes (128-byte aligned)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_in + plan.stages * stage_bytes);
    uint64_t* in_full = bars;                        // [SF_MAX_STAGES]
    uint64_t* in_empty = bars + SF_MAX_STAGES;       // [SF_MAX_STAGES]
    uint64_t* a_full = bars + 2 * SF_MAX_STAGES;     // [2]
    uint64_t* a_empty = a_full + 2;                  // [2]
    uint64_t* acc_full = a_full + 4;                 // [2]
    uint64_t* acc_empty = a_full + 6;                // [2]
    uint64_t* w_full = a_full + 8;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(a_full + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_items = plan.items > static_cast<int64_t>(blockIdx.x)
                             ? static_cast<int>((plan.items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (warp == SF_WARP_PROD && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_y);
        for (int s = 0; s < SF_MAX_STAGES; ++s) { mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], SF_DW_WARPS); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], SF_DW_WARPS); mbar_init(&a_empty[s], 1);
            mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], SF_EPI_WARPS);
        }
        mbar_init(w_full, 1);
        fence_mbar_init();
    }
    if (warp == SF_WARP_MMA) { tmem_alloc(tmem_holder, 512); tmem_relinquish(); }
    // operands of the bias MMA.  SWIZZLE_32B K-major: row r at r * 32, its 16-byte chunk q at ((q ^ ((r >> 2) & 1)) << 4);
    // only chunk 0 (k = 0 .. 7) is non-zero
    for (int r = threadIdx.x; r < 128 + n_out; r += SF_THREADS) {
        uint32_t first = 0x3F803F80u;                                  // (1.0, 1.0) in bf16
        uint8_t* row = s_ones + r * 32;
        if (r >= 128) {
            const float b = bias[r - 128];
            const __nv_bfloat16 hi = __float2bfloat16_rn(b);
            const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
            first = static_cast<uint32_t>(__bfloat16_as_ushort(hi)) | (static_cast<uint32_t>(__bfloat16_as_ushort(lo)) << 16);
            row = s_bias + (r - 128) * 32;
        }
        const int sw = (r >> 2) & 1;                                   // same for both tiles: both start 256-byte aligned
        *reinterpret_cast<uint4*>(row + ((0 ^ sw) << 4)) = make_uint4(first, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(row + ((1 ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    // item -> (image, strip, row block); row block fastest: a CTA's consecutive items continue down one strip, so the
    // two halo rows they share are L2 hits
    auto decode = [&](int64_t item, int& img, int& x0, int& y0) {
        const int rb = static_cast<int>(item % plan.rblocks);
        const int64_t r = item / plan.rblocks;
        const int strip = static_cast<int>(r % plan.strips);
        img = static_cast<int>(r / plan.strips);
        x0 = strip * 2 * plan.pairs;
        y0 = rb * SF_ROWS;
    };

    if (warp == SF_WARP_PROD) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_arrive_expect_tx(w_full, plan.cgroups * w_chunk);
            for (int g = 0; g < plan.cgroups; ++g) tma_load_2d(s_w + g * w_chunk, &tm_w, w_full, g * SF_CG, 0);
            int slot = 0;
            uint32_t phase = 0;
            for (int k = 0; k < my_items; ++k) {
                int img, x0, y0;
                decode(static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(k) * gridDim.x, img, x0, y0);
                for (int g = 0; g < plan.cgroups; ++g) {
                    mbar_wait_sleep(&in_empty[slot], phase ^ 1);
                    mbar_arrive_expect_tx(&in_full[slot], stage_bytes);
                    tma_load_4d(s_in + slot * stage_bytes, &tm_x, &in_full[slot], g * SF_CG, x0 - 1, y0 - 1, img);
                    if (++slot == plan.stages) { slot = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == SF_WARP_MMA) {
        // ================= MMA issuer =================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, static_cast<uint32_t>(n_out), 0, 0);
            const uint64_t desc = make_smem_desc(0, 0, 1024, SWZ_128B);
            const uint64_t a_f = (smem_u32(s_a) & 0x3FFFFu) >> 4, w_f = (smem_u32(s_w) & 0x3FFFFu) >> 4;
            const uint64_t ones_d = make_smem_desc(smem_u32(s_ones), 0, 256, SWZ_32B);
            const uint64_t bias_d = make_smem_desc(smem_u32(s_bias), 0, 256, SWZ_32B);
            mbar_wait(w_full, 0);
            int ab = 0;
            uint32_t a_phase = 0;
            for (int k = 0; k < my_items; ++k) {
                const int acc = k & 1;
                mbar_wait(&acc_empty[acc], ((k >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem + acc * 256;
                umma_f16_ss(d_tmem, ones_d, bias_d, idesc, 0u);          // accumulator = bias (hi + lo)
                for (int g = 0; g < plan.cgroups; ++g) {
                    mbar_wait_hot(&a_full[ab], a_phase);
                    tc_fence_after();
                    const uint64_t ad = desc | (a_f + ab * (SF_A_BYTES >> 4));
                    const uint64_t bd = desc | (w_f + g * (w_chunk >> 4));
#pragma unroll
                    for (int j = 0; j < 4; ++j) umma_f16_ss(d_tmem, ad + 2 * j, bd + 2 * j, idesc, 1u);
                    umma_commit(&a_empty[ab]);
                    if (++ab == 2) { ab = 0; a_phase ^= 1; }
                }
                umma_commit(&acc_full[acc]);
            }
        }
        __syncwarp();
    } else if (warp < SF_DW_WARPS) {
        // ================= depthwise 3x3 -> A tile =================
        const int tid = threadIdx.x;
        const int cq = tid & 15;                       // channel quad inside the group
        const int pair = tid >> 4;                     // column pair inside the strip
        const bool active = pair < plan.pairs;
        const int pair_c = active ? pair : 0;
        const int wpix = 2 * plan.pairs;               // pixels per tile row
        int slot = 0, ab = 0;
        uint32_t phase = 0, a_phase = 0;
        // fp32 pairs: FFMA2 (fma.rn.f32x2) does two of the 9 x 4 x 6 multiply-adds of a thread per issue slot — the
        // depthwise arithmetic is what paces the 128 -> 128 layer (r7a), and the FMA stream was 216 of its ~410 instructions
        f32x2_t wk[9][2];
        auto load_weights = [&](int g) {      // this thread's 4 channels of the 3x3 filters of channel group g
            const int ch = g * SF_CG + cq * 4;
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(dw + q * c + ch));
                wk[q][0] = f32x2_make(t.x, t.y);
                wk[q][1] = f32x2_make(t.z, t.w);
            }
        };
        if (plan.cgroups == 1) load_weights(0);       // one group: the filters stay in registers for the whole kernel
        for (int k = 0; k < my_items; ++k) {
            for (int g = 0; g < plan.cgroups; ++g) {
                if (plan.cgroups > 1) load_weights(g);
                f32x2_t acc[SF_ROWS][2][2];
#pragma unroll
                for (int r = 0; r < SF_ROWS; ++r)
#pragma unroll
                    for (int o = 0; o < 2; ++o) { acc[r][o][0] = 0ull; acc[r][o][1] = 0ull; }
                mbar_wait(&in_full[slot], phase);
                const uint32_t srow = smem_u32(s_in) + slot * stage_bytes + (2 * pair_c) * (SF_CG * 2) + cq * 8;
#pragma unroll
                for (int i = 0; i < SF_IN_ROWS; ++i) {
                    f32x2_t f[4][2];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint2 t;
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(t.x), "=r"(t.y) : "r"(srow + (i * in_w + q) * (SF_CG * 2)));
                        if (RELU_IN) {
                            const __nv_bfloat162 z = __float2bfloat162_rn(0.0f);
                            __nv_bfloat162 lo = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&t.x), z);
                            __nv_bfloat162 hi = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&t.y), z);
                            t.x = *reinterpret_cast<uint32_t*>(&lo);
                            t.y = *reinterpret_cast<uint32_t*>(&hi);
                        }
                        f[q][0] = f32x2_make_bits(t.x << 16, t.x & 0xffff0000u);      // channels 0, 1 as fp32
                        f[q][1] = f32x2_make_bits(t.y << 16, t.y & 0xffff0000u);      // channels 2, 3
                    }
                    // input row i is filter row ky of output row i - ky
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int r = i - ky;
                        if (r < 0 || r >= SF_ROWS) continue;
#pragma unroll
                        for (int o = 0; o < 2; ++o)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                f32x2_t a = acc[r][o][e];
                                a = f32x2_fma(f[o][e], wk[3 * ky][e], a);
                                a = f32x2_fma(f[o + 1][e], wk[3 * ky + 1][e], a);
                                a = f32x2_fma(f[o + 2][e], wk[3 * ky + 2][e], a);
                                acc[r][o][e] = a;
                            }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&in_empty[slot]);          // the slab has been read by this warp
                if (++slot == plan.stages) { slot = 0; phase ^= 1; }
                // A tile: row = pixel (r * wpix + column), 16-byte chunk (cq / 2) ^ (row & 7), 8 bytes per thread
                // the MMAs that read this buffer have retired.  (Backing off between polls — nanosleep instead of spinning
                // next to the epilogue warp that shares the scheduler — measured neutral: 0.813 / 1.164 / 0.457 vs
                // 0.814 / 1.211 / 0.441 ms on the three layers, r7b.)
                mbar_wait(&a_empty[ab], a_phase ^ 1);
                if (active) {
                    const uint32_t abase = smem_u32(s_a) + ab * SF_A_BYTES + (cq & 1) * 8;
#pragma unroll
                    for (int r = 0; r < SF_ROWS; ++r)
#pragma unroll
                        for (int o = 0; o < 2; ++o) {
                            const int p = r * wpix + 2 * pair_c + o;
                            const uint32_t addr = abase + p * 128 + (((cq >> 1) ^ (p & 7)) << 4);
                            float a0, a1, a2, a3;
                            f32x2_split(acc[r][o][0], a0, a1);
                            f32x2_split(acc[r][o][1], a2, a3);
                            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(pack_bf16x2(a0, a1)),
                                         "r"(pack_bf16x2(a2, a3))
                                         : "memory");
                        }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[ab]);
                if (++ab == 2) { ab = 0; a_phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue: lane = pixel =================
        const int quad = warp & 3;                      // TMEM lane quadrant of this warp (warps 12-15 -> 0-3)
        const int p = quad * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t slab0 = smem_u32(s_slab);
        const bool issuer = warp == SF_WARP_EPI0 && lane == 0;
        for (int k = 0; k < my_items; ++k) {
            int img, x0, y0;
            decode(static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(k) * gridDim.x, img, x0, y0);
            const int acc = k & 1;
            mbar_wait(&acc_full[acc], (k >> 1) & 1);
            tc_fence_after();
            // 128 output channels per round (n_out = 64: one round of 64): two TMEM loads in flight per wait, both slabs
            // staged, ONE proxy fence + barrier, then one TMA store per slab.  (One 64-channel group per round cost
            // ~1700 clk per group in fences / barrier latency and made the epilogue the pacing stage, r6y.)
            for (int n0 = 0; n0 < n_out; n0 += 128) {
                const int halves = n_out - n0 >= 128 ? 2 : 1;            // 64-channel groups in this round
                // both slabs were handed to the TMA unit in the previous round; the issuer waits until those stores have
                // been read out of shared memory, then everybody may overwrite them
                if (issuer) tma_store_wait_read0();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * SF_EPI_WARPS) : "memory");
#pragma unroll
                for (int gq = 0; gq < 2; ++gq) {
                    if (gq >= halves) break;
                    uint32_t r0[32], r1[32];
                    tmem_ld_32x32b_x32(tmem + lane_base + acc * 256 + n0 + gq * 64, r0);
                    tmem_ld_32x32b_x32(tmem + lane_base + acc * 256 + n0 + gq * 64 + 32, r1);
                    tmem_ld_wait();
                    if (n0 + 128 >= n_out && gq == halves - 1) {      // all TMEM reads of this accumulator are done
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[acc]);
                    }
                    uint32_t o[32];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        uint32_t v0 = pack_bf16x2(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
                        uint32_t v1 = pack_bf16x2(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
                        if (act == ISTVT_ACT_RELU) {       // on the rounded pair: max(round(x), 0) == round(max(x, 0))
                            const __nv_bfloat162 z = __float2bfloat162_rn(0.0f);
                            const __nv_bfloat162 m0 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&v0), z);
                            const __nv_bfloat162 m1 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&v1), z);
                            v0 = *reinterpret_cast<const uint32_t*>(&m0);
                            v1 = *reinterpret_cast<const uint32_t*>(&m1);
                        }
                        o[j] = v0;
                        o[16 + j] = v1;
                    }
                    const uint32_t srow = slab0 + gq * SF_SLAB_BYTES + p * 128;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        sts_u4(srow + ((q ^ (p & 7)) << 4), o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                }
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * SF_EPI_WARPS) : "memory");
                if (issuer) {
                    for (int gq = 0; gq < halves; ++gq) tma_store_4d(&tm_y, slab0 + gq * SF_SLAB_BYTES, n0 + gq * 64, x0, y0, img);
                    tma_store_commit();
                }
            }
        }
        if (issuer) tma_store_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == SF_WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

static int sepconv_fused_launch(const void* x, const float* dw, const void* pw, int64_t ld_pw, const float* bias, void* y,
                                int n, int h, int w, int c, int n_out, int relu_in, int act, cudaStream_t st) {
    SepFusedPlan pl{};
    pl.strips = (w + 2 * SF_MAX_PAIRS - 1) / (2 * SF_MAX_PAIRS);
    const int tw = (w + pl.strips - 1) / pl.strips;
    pl.pairs = (tw + 1) / 2;
    pl.rblocks = (h + SF_ROWS - 1) / SF_ROWS;
    pl.cgroups = c / SF_CG;
    pl.items = static_cast<int64_t>(n) * pl.strips * pl.rblocks;
    const int in_w = 2 * pl.pairs + 2;
    const int stage_bytes = SF_IN_ROWS * in_w * SF_CG * 2;
    const int fixed = 2 * SF_A_BYTES + 2 * SF_SLAB_BYTES + SF_ONES_BYTES + 256 * 32 + pl.cgroups * n_out * 128 +
                      1024 /*align*/ + 256 /*barriers*/;
    pl.stages = (227 * 1024 - fixed) / stage_bytes;
    if (pl.stages > SF_MAX_STAGES) pl.stages = SF_MAX_STAGES;
    if (pl.stages < 2) return ISTVT_ERR_UNSUPPORTED;
    const int smem = fixed + pl.stages * stage_bytes;

    CUtensorMap tm_x, tm_w, tm_y;
    {
        const uint64_t dims[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(w), static_cast<uint64_t>(h),
                                  static_cast<uint64_t>(n)};
        const uint64_t strides[3] = {static_cast<uint64_t>(c) * 2, static_cast<uint64_t>(w) * c * 2,
                                     static_cast<uint64_t>(h) * w * c * 2};
        const uint32_t box[4] = {SF_CG, static_cast<uint32_t>(in_w), SF_IN_ROWS, 1};
        int rc = encode_tmap(&tm_x, x, ISTVT_BF16, 4, dims, strides, box, 0);
        if (rc != ISTVT_OK) return rc;
    }
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(c), static_cast<uint64_t>(n_out)};
        const uint64_t strides[1] = {static_cast<uint64_t>(ld_pw) * 2};
        const uint32_t box[2] = {SF_CG, static_cast<uint32_t>(n_out)};
        int rc = encode_tmap(&tm_w, pw, ISTVT_BF16, 2, dims, strides, box, 3);
        if (rc != ISTVT_OK) return rc;
    }
    {
        const uint64_t dims[4] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(w), static_cast<uint64_t>(h),
                                  static_cast<uint64_t>(n)};
        const uint64_t strides[3] = {static_cast<uint64_t>(n_out) * 2, static_cast<uint64_t>(w) * n_out * 2,
                                     static_cast<uint64_t>(h) * w * n_out * 2};
        const uint32_t box[4] = {64, static_cast<uint32_t>(2 * pl.pairs), SF_ROWS, 1};
        int rc = encode_tmap(&tm_y, y, ISTVT_BF16, 4, dims, strides, box, 3);
        if (rc != ISTVT_OK) return rc;
    }
    const int64_t grid = pl.items < sm_count() ? pl.items : sm_count();
    if (relu_in) {
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(sepconv_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        sepconv_fused_kernel<true><<<static_cast<unsigned>(grid), SF_THREADS, smem, st>>>(tm_x, tm_w, tm_y, dw, bias, h, w, c,
                                                                                          n_out, act, pl);
    } else {
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(sepconv_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        sepconv_fused_kernel<false><<<static_cast<unsigned>(grid), SF_THREADS, smem, st>>>(tm_x, tm_w, tm_y, dw, bias, h, w, c,
                                                                                           n_out, act, pl);
    }
    count_launch();
    return launch_status();
}

}  // namespace istvt

using namespace istvt;

// y[n, h, w, n_out] (bf16) = act( pointwise( depthwise3x3_pad1( relu_in ? relu(x) : x ) ) + bias )
// x: bf16 NHWC [n, h, w, c]; dw: fp32 [3, 3, c]; pw: bf16 [n_out, c] (row pitch ld_pw, BatchNorm scale folded in);
// bias: fp32 [n_out].  c in {64, 128, 192, 256}, n_out a multiple of 64 up to 256, and the weights must fit next to the
// rings in shared memory (ISTVT_ERR_UNSUPPORTED otherwise: the caller runs istvt_dwconv3x3_fwd + istvt_gemm_fwd).
extern "C" int istvt_sepconv_fused_fwd(const void* x, const float* dw, const void* pw, int64_t ld_pw, const float* bias,
                                       void* y, int n, int h, int w, int c, int n_out, int relu_in, int act,
                                       istvt_stream_t stream) {
    ISTVT_REQUIRE(x && dw && pw && bias && y);
    ISTVT_REQUIRE(n > 0 && h > 0 && w > 0);
    ISTVT_REQUIRE(act == ISTVT_ACT_NONE || act == ISTVT_ACT_RELU);
    if (c % SF_CG != 0 || c > 256 || n_out % 64 != 0 || n_out < 64 || n_out > 256 || ld_pw < c || ld_pw % 8 != 0)
        return ISTVT_ERR_UNSUPPORTED;
    ISTVT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(pw) |
                    reinterpret_cast<uintptr_t>(dw) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0);
    return sepconv_fused_launch(x, dw, pw, ld_pw, bias, y, n, h, w, c, n_out, relu_in, act, static_cast<cudaStream_t>(stream));
}
