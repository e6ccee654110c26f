// Xception stem conv2 (3x3, stride 1, no padding, 32 -> 64 channels) + folded BatchNorm + ReLU (xception.py:122-123,
// 198-200) on the strip pipeline of sepconv_fused.cu.
//
// The 9-tap TMA formulation (gemm_tcgen05.cu, conv mode) runs at the TMA unit's box-ROW rate: with 32 input channels a
// pixel is a 64-byte row and a 128-pixel tile needs 9 x 128 of them (0.21 of the HBM roofline, profiles/README.md
// r4d / r4e); gathering the operand with per-thread global loads (conv3x3_tc.cu) was slower still (r6l: dependent loads).
// Here the input reaches shared memory ONCE per item as a 5-row TMA slab (200 box rows for 114 output pixels), and the
// im2col is a shared-memory to shared-memory rearrangement:
//   * item = (image, strip of <= 38 output columns, 3 output rows) = <= 114 pixels = TMEM lanes (as sepconv_fused.cu);
//   * for filter row ky the three taps of output pixel (r, x) are 3 x 64 CONTIGUOUS bytes of slab row r + ky: 4 gather warps
//     (thread = pixel) copy them as 12 16-byte chunks into two K-major operand tiles — k = 0..63 (kx 0, 1) as a
//     SWIZZLE_128B row, k = 64..95 (kx 2) as a SWIZZLE_64B row;
//   * 6 tcgen05.mma (M 128, N 64, K 16) per filter row against weights resident in shared memory in the same two-tile
//     form; the bias enters through a K = 16 MMA of ones x (bias hi, bias lo); accumulator double buffered in TMEM;
//   * 4 epilogue warps: TMEM -> ReLU -> bf16 -> swizzled slab -> one 4-D TMA store (clipped at the image border).
// MEASURED (profiles/README.md r7h): parity green, but 1.01 ms against the 9-tap path's 0.91 ms at the C2 size — the
// shared-memory im2col reads and writes 66 KB per 114-pixel item (9x the input) and the MMAs read the 72 KB back, so the
// kernel sits on the shared-memory port (~2200 clk per item) instead of on the TMA row rate.  Kept as the opt-in
// alternative ISTVT_CONV2_KERNEL=strip; the default for conv2 stays the 9-tap TMA formulation.
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace istvt {

constexpr int CS_CI = 32, CS_CO = 64;
constexpr int CS_ROWS = 3, CS_IN_ROWS = 5;
constexpr int CS_MAX_COLS = 38;                  // output columns per strip: 38 x 3 = 114 lanes
constexpr int CS_GATHER_WARPS = 4, CS_EPI_WARPS = 4;
constexpr int CS_WARP_PROD = 8, CS_WARP_MMA = 9;                       // warps 0-3 gather, 4-7 epilogue (quadrant = warp & 3)
constexpr int CS_THREADS = 32 * 10;
constexpr int CS_T0_BYTES = 128 * 128;           // [128 px x 64 k] SW128
constexpr int CS_T1_BYTES = 128 * 64;            // [128 px x 32 k] SW64
constexpr int CS_A_KY = CS_T0_BYTES + CS_T1_BYTES;                      // 24 KB per filter row
constexpr int CS_W0_BYTES = CS_CO * 128, CS_W1_BYTES = CS_CO * 64;      // 8 KB + 4 KB per filter row
constexpr int CS_W_KY = CS_W0_BYTES + CS_W1_BYTES;
constexpr int CS_SLAB_BYTES = 128 * 128;         // output staging: 128 px x 64 channels bf16
constexpr int CS_ONES_BYTES = 128 * 32, CS_BIAS_BYTES = CS_CO * 32;
constexpr int CS_STAGES = 4;

struct ConvStripPlan {
    int cols, strips, rblocks;
    int64_t items;
};

__device__ __forceinline__ void cs_tma_store_4d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

__global__ void __launch_bounds__(CS_THREADS, 1)
conv3x3_c32_strip_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                         const __nv_bfloat16* __restrict__ wt, const float* __restrict__ bias, int act,
                         const ConvStripPlan plan) {
    extern __shared__ uint8_t cs_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(cs_raw) + 1023) & ~uintptr_t(1023));
    const int in_w = plan.cols + 2;
    const uint32_t stage_bytes = static_cast<uint32_t>(CS_IN_ROWS * in_w * CS_CI * 2);
    uint8_t* s_a = smem;                                   // 3 x (16 KB + 8 KB)
    uint8_t* s_w = s_a + 3 * CS_A_KY;                      // 3 x (8 KB + 4 KB)
    uint8_t* s_slab = s_w + 3 * CS_W_KY;                   // 16 KB
    uint8_t* s_ones = s_slab + CS_SLAB_BYTES;              // 4 KB
    uint8_t* s_bias = s_ones + CS_ONES_BYTES;              // 2 KB
    uint8_t* s_in = s_bias + CS_BIAS_BYTES;                // CS_STAGES x stage_bytes
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_in + CS_STAGES * ((stage_bytes + 127) & ~127u));
    uint64_t* in_full = bars;                  // [CS_STAGES]
    uint64_t* in_empty = bars + CS_STAGES;     // [CS_STAGES]
    uint64_t* a_full = bars + 2 * CS_STAGES;   // [3] per filter row
    uint64_t* a_empty = a_full + 3;            // [3]
    uint64_t* acc_full = a_full + 6;           // [2]
    uint64_t* acc_empty = a_full + 8;          // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(a_full + 10);
    const uint32_t stage_pitch = (stage_bytes + 127) & ~127u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_items = plan.items > static_cast<int64_t>(blockIdx.x)
                             ? static_cast<int>((plan.items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (warp == CS_WARP_PROD && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_y);
        for (int s = 0; s < CS_STAGES; ++s) { mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], CS_GATHER_WARPS); }
        for (int s = 0; s < 3; ++s) { mbar_init(&a_full[s], CS_GATHER_WARPS); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], CS_EPI_WARPS); }
        fence_mbar_init();
    }
    if (warp == CS_WARP_MMA) { tmem_alloc(tmem_holder, 128); tmem_relinquish(); }
    // weights [co][ky][kx][ci] -> per filter row: W0 = row co, k = kx * 32 + ci for kx 0, 1 (128 B, SW128: 16-byte chunk q at
    // (q ^ (co & 7)) << 4) and W1 = row co, k = ci of kx 2 (64 B, SW64: chunk q at (q ^ ((co >> 1) & 3)) << 4)
    for (int i = threadIdx.x; i < 3 * CS_CO * 12; i += CS_THREADS) {
        const int q = i % 12, co = (i / 12) % CS_CO, ky = i / (12 * CS_CO);
        const uint4 v = *reinterpret_cast<const uint4*>(wt + ((co * 3 + ky) * 3) * CS_CI + q * 8);
        uint8_t* base = s_w + ky * CS_W_KY;
        if (q < 8) *reinterpret_cast<uint4*>(base + co * 128 + ((q ^ (co & 7)) << 4)) = v;
        else       *reinterpret_cast<uint4*>(base + CS_W0_BYTES + co * 64 + (((q - 8) ^ ((co >> 1) & 3)) << 4)) = v;
    }
    // operands of the bias MMA (SW32 K-major, see sepconv_fused.cu)
    for (int r = threadIdx.x; r < 128 + CS_CO; r += CS_THREADS) {
        uint32_t first = 0x3F803F80u;
        uint8_t* row = s_ones + r * 32;
        if (r >= 128) {
            const float b = bias[r - 128];
            const __nv_bfloat16 hi = __float2bfloat16_rn(b);
            const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
            first = static_cast<uint32_t>(__bfloat16_as_ushort(hi)) | (static_cast<uint32_t>(__bfloat16_as_ushort(lo)) << 16);
            row = s_bias + (r - 128) * 32;
        }
        const int sw = (r >> 2) & 1;
        *reinterpret_cast<uint4*>(row + ((0 ^ sw) << 4)) = make_uint4(first, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(row + ((1 ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    auto decode = [&](int64_t item, int& img, int& x0, int& y0) {
        const int rb = static_cast<int>(item % plan.rblocks);
        const int64_t r = item / plan.rblocks;
        const int strip = static_cast<int>(r % plan.strips);
        img = static_cast<int>(r / plan.strips);
        x0 = strip * plan.cols;
        y0 = rb * CS_ROWS;
    };

    if (warp == CS_WARP_PROD) {
        // ================= TMA producer: one 5-row slab per item =================
        if (lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            for (int k = 0; k < my_items; ++k) {
                int img, x0, y0;
                decode(static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(k) * gridDim.x, img, x0, y0);
                mbar_wait_sleep(&in_empty[slot], phase ^ 1);
                mbar_arrive_expect_tx(&in_full[slot], stage_bytes);
                tma_load_4d(s_in + slot * stage_pitch, &tm_x, &in_full[slot], 0, x0, y0, img);
                if (++slot == CS_STAGES) { slot = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == CS_WARP_MMA) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, CS_CO, 0, 0);
            const uint64_t d128 = make_smem_desc(0, 0, 1024, SWZ_128B), d64 = make_smem_desc(0, 0, 512, SWZ_64B);
            const uint64_t a_f = (smem_u32(s_a) & 0x3FFFFu) >> 4, w_f = (smem_u32(s_w) & 0x3FFFFu) >> 4;
            const uint64_t ones_d = make_smem_desc(smem_u32(s_ones), 0, 256, SWZ_32B);
            const uint64_t bias_d = make_smem_desc(smem_u32(s_bias), 0, 256, SWZ_32B);
            for (int k = 0; k < my_items; ++k) {
                const int acc = k & 1;
                mbar_wait(&acc_empty[acc], ((k >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem + acc * CS_CO;
                umma_f16_ss(d_tmem, ones_d, bias_d, idesc, 0u);                    // accumulator = bias
                for (int ky = 0; ky < 3; ++ky) {
                    mbar_wait_hot(&a_full[ky], k & 1);
                    tc_fence_after();
                    const uint64_t a0 = d128 | (a_f + ky * (CS_A_KY >> 4)), b0 = d128 | (w_f + ky * (CS_W_KY >> 4));
                    const uint64_t a1 = d64 | (a_f + (ky * CS_A_KY + CS_T0_BYTES >> 4)),
                                   b1 = d64 | (w_f + (ky * CS_W_KY + CS_W0_BYTES >> 4));
#pragma unroll
                    for (int j = 0; j < 4; ++j) umma_f16_ss(d_tmem, a0 + 2 * j, b0 + 2 * j, idesc, 1u);     // kx 0, 1
#pragma unroll
                    for (int j = 0; j < 2; ++j) umma_f16_ss(d_tmem, a1 + 2 * j, b1 + 2 * j, idesc, 1u);     // kx 2
                    umma_commit(&a_empty[ky]);
                }
                umma_commit(&acc_full[acc]);
            }
        }
        __syncwarp();
    } else if (warp < CS_GATHER_WARPS) {
        // ================= im2col: slab rows -> A tiles, thread = output pixel =================
        const int p = threadIdx.x;                         // 0 .. 127
        const bool active = p < CS_ROWS * plan.cols;
        const int r = active ? p / plan.cols : 0, col = active ? p - r * plan.cols : 0;
        const int sw128 = p & 7, sw64 = (p >> 1) & 3;
        int slot = 0;
        uint32_t phase = 0;
        for (int k = 0; k < my_items; ++k) {
            mbar_wait(&in_full[slot], phase);
            const uint32_t src0 = smem_u32(s_in) + slot * stage_pitch + (r * in_w + col) * (CS_CI * 2);
            for (int ky = 0; ky < 3; ++ky) {
                uint4 v[12];
                if (active) {
#pragma unroll
                    for (int q = 0; q < 12; ++q) v[q] = lds_u4(src0 + ky * in_w * (CS_CI * 2) + q * 16);
                }
                mbar_wait(&a_empty[ky], (k & 1) ^ 1);      // the previous item's MMAs of this filter row have retired
                if (active) {
                    const uint32_t t0 = smem_u32(s_a) + ky * CS_A_KY + p * 128;
                    const uint32_t t1 = smem_u32(s_a) + ky * CS_A_KY + CS_T0_BYTES + p * 64;
#pragma unroll
                    for (int q = 0; q < 8; ++q) sts_u4(t0 + ((q ^ sw128) << 4), v[q].x, v[q].y, v[q].z, v[q].w);
#pragma unroll
                    for (int q = 0; q < 4; ++q) sts_u4(t1 + ((q ^ sw64) << 4), v[8 + q].x, v[8 + q].y, v[8 + q].z, v[8 + q].w);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[ky]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&in_empty[slot]);
            if (++slot == CS_STAGES) { slot = 0; phase ^= 1; }
        }
    } else if (warp < CS_GATHER_WARPS + CS_EPI_WARPS) {
        // ================= epilogue: lane = pixel =================
        const int quad = warp & 3;
        const int p = quad * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t slab = smem_u32(s_slab);
        const bool issuer = warp == CS_GATHER_WARPS && lane == 0;
        for (int k = 0; k < my_items; ++k) {
            int img, x0, y0;
            decode(static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(k) * gridDim.x, img, x0, y0);
            const int acc = k & 1;
            mbar_wait(&acc_full[acc], (k >> 1) & 1);
            tc_fence_after();
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(tmem + lane_base + acc * CS_CO, r0);
            tmem_ld_32x32b_x32(tmem + lane_base + acc * CS_CO + 32, r1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
            uint32_t o[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t v0 = pack_bf16x2(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
                uint32_t v1 = pack_bf16x2(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
                if (act == ISTVT_ACT_RELU) {
                    const __nv_bfloat162 z = __float2bfloat162_rn(0.0f);
                    const __nv_bfloat162 m0 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&v0), z);
                    const __nv_bfloat162 m1 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&v1), z);
                    v0 = *reinterpret_cast<const uint32_t*>(&m0);
                    v1 = *reinterpret_cast<const uint32_t*>(&m1);
                }
                o[j] = v0;
                o[16 + j] = v1;
            }
            if (issuer) tma_store_wait_read0();            // the previous item's store has left the slab
            asm volatile("bar.sync 1, %0;" ::"n"(32 * CS_EPI_WARPS) : "memory");
            const uint32_t srow = slab + p * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q) sts_u4(srow + ((q ^ (p & 7)) << 4), o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * CS_EPI_WARPS) : "memory");
            if (issuer) {
                cs_tma_store_4d(&tm_y, slab, 0, x0, y0, img);
                tma_store_commit();
            }
        }
        if (issuer) tma_store_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == CS_WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem, 128);
    }
}

// host: conv2-shaped 3x3 convolution (cin 32, cout 64, bf16, pad 0) on the strip kernel
int conv3x3_c32_strip_launch(const void* x, const void* wt, const float* bias, void* y, int n, int h, int w, int act,
                             cudaStream_t st) {
    const int ho = h - 2, wo = w - 2;
    ConvStripPlan pl{};
    pl.strips = (wo + CS_MAX_COLS - 1) / CS_MAX_COLS;
    pl.cols = (wo + pl.strips - 1) / pl.strips;
    pl.rblocks = (ho + CS_ROWS - 1) / CS_ROWS;
    pl.items = static_cast<int64_t>(n) * pl.strips * pl.rblocks;
    const int in_w = pl.cols + 2;
    const int stage_bytes = CS_IN_ROWS * in_w * CS_CI * 2;
    const int smem = 3 * CS_A_KY + 3 * CS_W_KY + CS_SLAB_BYTES + CS_ONES_BYTES + CS_BIAS_BYTES +
                     CS_STAGES * ((stage_bytes + 127) & ~127) + 1024 + 256;
    CUtensorMap tm_x, tm_y;
    {
        const uint64_t dims[4] = {CS_CI, static_cast<uint64_t>(w), static_cast<uint64_t>(h), static_cast<uint64_t>(n)};
        const uint64_t strides[3] = {CS_CI * 2, static_cast<uint64_t>(w) * CS_CI * 2, static_cast<uint64_t>(h) * w * CS_CI * 2};
        const uint32_t box[4] = {CS_CI, static_cast<uint32_t>(in_w), CS_IN_ROWS, 1};
        int rc = encode_tmap(&tm_x, x, ISTVT_BF16, 4, dims, strides, box, 0);
        if (rc != ISTVT_OK) return rc;
    }
    {
        const uint64_t dims[4] = {CS_CO, static_cast<uint64_t>(wo), static_cast<uint64_t>(ho), static_cast<uint64_t>(n)};
        const uint64_t strides[3] = {CS_CO * 2, static_cast<uint64_t>(wo) * CS_CO * 2, static_cast<uint64_t>(ho) * wo * CS_CO * 2};
        const uint32_t box[4] = {CS_CO, static_cast<uint32_t>(pl.cols), CS_ROWS, 1};
        int rc = encode_tmap(&tm_y, y, ISTVT_BF16, 4, dims, strides, box, 3);
        if (rc != ISTVT_OK) return rc;
    }
    const int64_t grid = pl.items < sm_count() ? pl.items : sm_count();
    ISTVT_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_c32_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    conv3x3_c32_strip_kernel<<<static_cast<unsigned>(grid), CS_THREADS, smem, st>>>(
        tm_x, tm_y, static_cast<const __nv_bfloat16*>(wt), bias, act, pl);
    count_launch();
    return launch_status();
}

}  // namespace istvt
