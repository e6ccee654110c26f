// tcgen05 / TMEM / TMA GEMM for sm_100a:  C = act(A · Wᵀ + bias) + residual
//
//   A  : bf16 [M, K]  row-major (K-major operand)      — token rows / NHWC pixels
//   W  : bf16 [N, K]  row-major (K-major operand)      — nn.Linear / 1x1-conv weight layout
//   C  : bf16 or fp32 [M, N]
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (single lane issues
// tcgen05.mma), warp 2 = TMEM allocator, warps 4.. = epilogue (TMEM -> registers -> swizzled smem transpose -> coalesced global; 32 rows x 64 columns per warp).
// Three pipelines: smem ring (full/empty mbarriers, TMA <-> MMA), TMEM double buffer
// (tmem_full/tmem_empty, MMA <-> epilogue), and the static tile schedule (tile = blockIdx.x + i*grid),
// n-fastest so that the CTAs in flight share A rows through L2 and the small W stays L2-resident.
//
// The same kernel runs the dense 3x3 convolution of the Xception stem as an implicit GEMM: in "conv
// mode" every k-block belongs to a filter tap (ky, kx) and the A tile of that tap is the same 2-D box
// of the NHWC activation matrix shifted down by ky*W_in + kx pixel rows (outputs are computed on the
// input's W_in-wide grid; the epilogue drops the two junk columns/rows and compacts the row index).
#include "gemm_common.cuh"

#include <stdlib.h>

namespace istvt {


template <int BN, int BK, bool PLAIN>
struct GemmCfg {
    // epilogue warp = lane quadrant x column group.  bf16 output: 32-column groups (every GEMM on this kernel has
    // K <= 256 or is the 9-tap conv: the tile period was the epilogue's latency chain — 2.5 us per 128 x 128 tile with
    // 64-column warps, profiles/README.md r3k — so the tile is spread over twice the warps); fp32 / residual: 64.
    static constexpr int WARP_COLS = PLAIN ? 32 : EPI_COLS;
    static constexpr int EPI_WARPS = 4 * (BN / WARP_COLS);
    static constexpr int SLAB = PLAIN ? EPI_SLAB_PLAIN_BYTES : EPI_SLAB_BYTES;
    static constexpr int THREADS = 128 + EPI_WARPS * 32;
    static constexpr int A_BYTES = GEMM_BLOCK_M * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BUDGET = 192 * 1024;
    static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * SLAB + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr uint32_t SWZ = (BK == 64) ? SWZ_128B : SWZ_64B;
    static constexpr uint32_t SBO = 8 * BK * 2;  // bytes between 8-row groups of a K-major swizzled tile
};

template <int BN, int BK, bool PLAIN_BF16>
__global__ void __launch_bounds__(GemmCfg<BN, BK, PLAIN_BF16>::THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const GemmParams p) {
    using Cfg = GemmCfg<BN, BK, PLAIN_BF16>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    uint8_t* smem_bres = smem;          // b_resident: [num_kb][B_BYTES] in front of a shorter A ring
    uint8_t* smem_epi = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::EPI_WARPS * Cfg::SLAB);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    uint64_t* b_full = bars + 2 * STAGES + 5;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int kb_per_tap = (p.K + BK - 1) / BK;
    const int num_kb = kb_per_tap * p.taps;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int64_t m_tiles = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
    const int64_t total_tiles = m_tiles * n_tiles;
    // Weights resident in shared memory (conv2 on pixel pairs: 7 x 16 KB): without it every k-block fetches its B box from
    // L2 again, and A + B fills ran at ~3/4 of one SM's L2 port (profiles/README.md r8h).  The A ring shrinks accordingly.
    const bool b_res = p.b_resident != 0;
    int nst = STAGES;
    if (b_res) {
        nst = (STAGES * Cfg::STAGE_BYTES - num_kb * Cfg::B_BYTES) / Cfg::A_BYTES;
        if (nst > STAGES) nst = STAGES;
        smem_a = smem + num_kb * Cfg::B_BYTES;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], p.split_producer ? 2 : 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], Cfg::EPI_WARPS);
        }
        mbar_init(b_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================== TMA producer (one elected thread) =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            if (b_res && total_tiles > static_cast<int64_t>(blockIdx.x)) {
                mbar_arrive_expect_tx(b_full, static_cast<uint32_t>(num_kb) * Cfg::B_BYTES);
                int kb = 0;
                for (int tap = 0; tap < p.taps; ++tap)
                    for (int kcol = 0; kcol < p.K; kcol += BK, ++kb)
                        tma_load_2d(smem_bres + kb * Cfg::B_BYTES, &tm_b, b_full, tap * p.K + kcol, 0);
            }
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int64_t m_blk = tile / n_tiles;
                const int n_blk = static_cast<int>(tile % n_tiles);
                const int64_t m0 = m_blk * GEMM_BLOCK_M;
                const int n0 = n_blk * BN;
                // taps x k-blocks as nested counters: this single thread issues every TMA of the CTA, and for the 9-tap
                // conv (one 12 KB stage per tap) its instruction count per stage — not the ring, not the MMAs — set the
                // tile period (r4a: a second producer thread for the B boxes alone gave 10 %)
                int bcol = 0;
                for (int tap = 0; tap < p.taps; ++tap) {
                    const int row = static_cast<int>(m0) + p.tap_shift[tap];     // 0 for a plain GEMM
                    for (int kcol = 0; kcol < p.K; kcol += BK) {
                        mbar_wait_hot(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], (p.split_producer || b_res) ? Cfg::A_BYTES : Cfg::STAGE_BYTES);
                        tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tm_a, &full_bar[stage], kcol, row);
                        if (!p.split_producer && !b_res)
                            tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tm_b, &full_bar[stage], bcol + kcol, n0);
                        if (++stage == nst) { stage = 0; phase ^= 1; }
                    }
                    bcol += p.K;
                }
            }
        }
        __syncwarp();
    } else if (warp == 3 && p.split_producer) {
        // ===================== second TMA producer: the B (weight) boxes =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n0 = static_cast<int>(tile % n_tiles) * BN;
                int bcol = 0;
                for (int tap = 0; tap < p.taps; ++tap) {
                    for (int kcol = 0; kcol < p.K; kcol += BK) {
                        mbar_wait_hot(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::B_BYTES);
                        tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tm_b, &full_bar[stage], bcol + kcol, n0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    bcol += p.K;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (ONE elected thread; see gemm_tcgen05_2cta.cu for why it is this lean) =====
        if (elect_one()) {
            constexpr int KSTEPS = BK / UMMA_K;
            const uint64_t desc_hi = make_smem_desc(0, 0, Cfg::SBO, Cfg::SWZ);
            const uint32_t a_field0 = (smem_u32(smem_a) & 0x3FFFFu) >> 4;
            const uint32_t b_field0 = (smem_u32(b_res ? smem_bres : smem_b) & 0x3FFFFu) >> 4;
            if (b_res) {
                mbar_wait(b_full, 0);
                tc_fence_after();
            }
            // MMAs of the last k-block of a tap (ragged K: the TMA zero-fills, whole zero k-steps are skipped)
            const int last_steps = (p.K - (kb_per_tap - 1) * BK + UMMA_K - 1) / UMMA_K;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_blk = static_cast<int>(tile % n_tiles);
                int n_eff = p.N - n_blk * BN;
                n_eff = n_eff >= BN ? BN : ((n_eff + 15) & ~15);
                const uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, static_cast<uint32_t>(n_eff), 0, 0);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                int kb_in_tap = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_hot(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = desc_hi | (a_field0 + stage * (Cfg::A_BYTES >> 4));
                    const uint64_t b_desc = desc_hi | (b_field0 + (b_res ? kb : stage) * (Cfg::B_BYTES >> 4));
                    const bool tap_end = (++kb_in_tap == kb_per_tap);
                    if (tap_end) kb_in_tap = 0;
                    if (!tap_end || last_steps == KSTEPS) {
#pragma unroll
                        for (int k = 0; k < KSTEPS; ++k)
                            umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (k != 0 || kb != 0) ? 1u : 0u);
                    } else {
                        for (int k = 0; k < last_steps; ++k)
                            umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (k != 0 || kb != 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);                      // smem slot free when these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;
        const int quad = warp & 3;          // TMEM lane quadrant this warp may access
        const int col0 = (ew >> 2) * Cfg::WARP_COLS;
        const uint32_t slab = smem_u32(smem_epi + ew * Cfg::SLAB);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int64_t m_blk = tile / n_tiles;
            const int n_blk = static_cast<int>(tile % n_tiles);
            const int64_t m = m_blk * GEMM_BLOCK_M + quad * 32 + lane;
            // destination row (conv mode compacts the padded grid)
            bool row_ok = m < p.M;
            int64_t drow = m;
            int ccol = n_blk * BN + col0;
            if (p.taps != 1) {
                // pixel-pair mode: accumulator row m holds pixels 2m (columns 0..63) and 2m + 1 (columns 64..127)
                int64_t lin = m;
                if (p.pair) { lin = 2 * m + (col0 >> 6); ccol = col0 & 63; }
                const int64_t img_sz = (int64_t)p.conv_h_in * p.conv_w_in;
                const int64_t img = lin / img_sz;
                const int rem = static_cast<int>(lin - img * img_sz);
                const int y = rem / p.conv_w_in;
                const int x = rem - y * p.conv_w_in;
                const int ho = p.conv_h_in - 2, wo = p.conv_w_in - 2;
                row_ok = row_ok && (y < ho) && (x < wo);
                drow = (img * ho + y) * wo + x;
            }
            int drow_t[8];
            const int drow_lane = row_ok ? static_cast<int>(drow) : -1;
            epilogue_rows(drow_lane, lane, drow_t);
            if constexpr (!PLAIN_BF16) epilogue_prefetch_residual(p, drow_lane, n_blk * BN + col0, Cfg::WARP_COLS);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            uint64_t* te = &tmem_empty[acc];
            gemm_epilogue_64<PLAIN_BF16, PLAIN_BF16 ? 1 : 2>(p, tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + col0, slab, drow_lane, drow_t,
                             ccol, lane, [&]() {
                                 tc_fence_before();
                                 __syncwarp();
                                 if (lane == 0) mbar_arrive(te);
                             });
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int BN, int BK>
static int launch_gemm(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const GemmParams& p, cudaStream_t stream) {
    const int n_tiles = (p.N + BN - 1) / BN;
    const int64_t m_tiles = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
    const int64_t total = m_tiles * n_tiles;
    int grid = sm_count();
    if (total < grid) grid = static_cast<int>(total);
    // 9-tap conv: A and B boxes issued by two threads (warp 0 / warp 3); ISTVT_G1_SPLIT_PRODUCER=0/1 forces it off / on
    static const int split_env = []() { const char* e = getenv("ISTVT_G1_SPLIT_PRODUCER"); return e ? atoi(e) : -1; }();
    GemmParams q = p;
    q.split_producer = split_env >= 0 ? split_env : (p.taps > 1 ? 1 : 0);
    if (q.b_resident) {   // one N tile, and room for the weights plus at least two A stages
        using C0 = GemmCfg<BN, BK, true>;
        const int num_kb = ((p.K + BK - 1) / BK) * p.taps;
        if (n_tiles != 1 || num_kb * C0::B_BYTES + 2 * C0::A_BYTES > C0::STAGES * C0::STAGE_BYTES) q.b_resident = 0;
        else q.split_producer = 0;
    }
    if (!p.c_f32 && p.residual == nullptr) {
        using Cfg = GemmCfg<BN, BK, true>;
        auto kern = gemm_tcgen05_kernel<BN, BK, true>;
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tm_a, tm_b, q);
    } else {
        using Cfg = GemmCfg<BN, BK, false>;
        auto kern = gemm_tcgen05_kernel<BN, BK, false>;
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tm_a, tm_b, q);
    }
    count_launch();
    return launch_status();
}

// Shared host path of istvt_gemm_fwd / istvt_conv3x3_fwd (bf16).
int gemm_bf16_dispatch(const void* a, int64_t a_rows, int64_t lda, const void* w, int64_t ldw, int w_cols,
                       const GemmParams& p, cudaStream_t stream) {
    ISTVT_REQUIRE(a && w && p.C);
    ISTVT_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0);
    ISTVT_REQUIRE(p.N % 8 == 0 && p.K % 8 == 0);
    ISTVT_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && p.ldc % 8 == 0);
    ISTVT_REQUIRE(p.residual == nullptr || p.ldr % 4 == 0);
    ISTVT_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    ISTVT_REQUIRE(p.M < (int64_t(1) << 31) - 4096);

    // Plain GEMMs with N >= 256 (every transformer Linear, block-3 pointwise/skip) run on CTA pairs.
    if (p.taps == 1 && p.N >= 256) return launch_gemm_2cta(a, lda, w, ldw, p, stream);

    // tile configuration: BK = 64 (128B swizzle) unless the per-tap K is 32; BN by N.
    const int bk = (p.K % 64 == 0 || p.K > 64) ? 64 : 32;
    ISTVT_REQUIRE(bk == 64 || p.K == 32);
    const int bn = p.N <= 64 ? 64 : 128;

    CUtensorMap tm_a, tm_b;
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(a_rows)};
        const uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
        const uint32_t box[2] = {static_cast<uint32_t>(bk), GEMM_BLOCK_M};
        int rc = encode_tmap(&tm_a, a, ISTVT_BF16, 2, dims, strides, box, bk == 64 ? 3 : 2);
        if (rc != ISTVT_OK) return rc;
    }
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(w_cols), static_cast<uint64_t>(p.N)};
        const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
        const uint32_t box[2] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(bn)};
        int rc = encode_tmap(&tm_b, w, ISTVT_BF16, 2, dims, strides, box, bk == 64 ? 3 : 2);
        if (rc != ISTVT_OK) return rc;
    }
    if (bk == 64) {
        if (bn == 64) return launch_gemm<64, 64>(tm_a, tm_b, p, stream);
        return launch_gemm<128, 64>(tm_a, tm_b, p, stream);
    } else {
        if (bn == 64) return launch_gemm<64, 32>(tm_a, tm_b, p, stream);
        return launch_gemm<128, 32>(tm_a, tm_b, p, stream);
    }
}

// conv3x3_tc.cu / conv3x3_strip.cu: conv2's shape (32 -> 64 channels) on dedicated kernels
int conv3x3_c32_tc_launch(const void* x, const void* wt, const float* bias, void* y, int n, int h, int w, int act,
                          cudaStream_t st);
int conv3x3_c32_strip_launch(const void* x, const void* wt, const float* bias, void* y, int n, int h, int w, int act,
                             cudaStream_t st);

// bf16 dense 3x3 stride-1 pad-0 convolution (conv2 of the stem) on the implicit-GEMM path.
int conv3x3_bf16(const void* x, const void* wt, const float* bias, void* y, int n, int h, int w, int cin, int cout,
                 int act, cudaStream_t stream) {
    // conv2's shape (32 -> 64): ISTVT_CONV2_KERNEL = taps (default: the 9-tap TMA formulation below, the path of every
    // other shape; 0.91 ms at the C2 size) | strip (TMA slab -> shared-memory im2col -> tcgen05, conv3x3_strip.cu: 1.01 ms,
    // r7h — the im2col moves 9x the input through the shared-memory port) | gather (per-thread global loads,
    // conv3x3_tc.cu: 1.23 ms, r6l).  Read per call so that a check can exercise all three in one process.
    const char* ke = getenv("ISTVT_CONV2_KERNEL");
    const char* legacy = getenv("ISTVT_CONV2_TC");
    const bool shape_ok = cin == 32 && cout == 64 && (act == ISTVT_ACT_NONE || act == ISTVT_ACT_RELU) && h >= 3 && w >= 3 &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wt) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (shape_ok) {
        const bool gather = (ke != nullptr && ke[0] == 'g') || (ke == nullptr && legacy != nullptr && atoi(legacy) != 0);
        const bool strip = ke != nullptr && ke[0] == 's';
        if (gather) return conv3x3_c32_tc_launch(x, wt, bias, y, n, h, w, act, stream);
        if (strip) return conv3x3_c32_strip_launch(x, wt, bias, y, n, h, w, act, stream);
    }
    GemmParams p{};
    p.M = static_cast<int64_t>(n) * h * w;  // outputs on the input's grid; junk rows/cols dropped in the epilogue
    p.N = cout; p.K = cin;
    p.taps = 9; p.conv_w_in = w; p.conv_h_in = h;
    for (int t = 0; t < 9; ++t) p.tap_shift[t] = (t / 3) * w + (t % 3);
    p.C = y; p.ldc = cout;
    p.bias = bias; p.residual = nullptr; p.ldr = 0;
    p.act = act; p.c_f32 = 0;
    return gemm_bf16_dispatch(x, p.M, cin, wt, 9 * static_cast<int64_t>(cin), 9 * cin, p, stream);
}

// ------------------------------------------------------------------------------------------
// conv2 (32 -> 64 channels) on PIXEL PAIRS.  The 9-tap formulation above fetches 64-byte box rows (32 channels), 9 x 128 of
// them per 128-pixel tile, and is bound by the TMA unit's row rate (0.21-0.27 of the HBM roofline).  Two
// adjacent NHWC pixels are 128 contiguous bytes, so the activation is ALSO a matrix of pixel pairs [n h w / 2, 64]: the
// output pair (i, i + 1) of the input's linear pixel grid needs the input pixels i + ky W + {0..3} = two aligned pairs
// when ky W is even, the middle of three pairs when it is odd.  That makes 6 (W even) or 7 (W odd) k-blocks of 64 with
// 128-byte rows for TWO output pixels — 2.6 x fewer TMA rows per pixel — against a weight matrix [2 x 64, taps x 64]
// that holds W[co, ky, kx, ci] at (pixel p, co) x (tap, pixel j, ci) with kx = 2 jblk + j - s - p and zeros elsewhere
// (1.3-1.6 x the multiply-adds, on a tensor pipe that has them to spare).  Same kernel, same epilogue.
// ------------------------------------------------------------------------------------------
__global__ void conv3x3_pair_pack_kernel(const __nv_bfloat16* __restrict__ wt, __nv_bfloat16* __restrict__ wpair,
                                         int parity, int taps) {
    const int kcols = taps * 64;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 128 * kcols) return;
    const int r = idx / kcols, c = idx - r * kcols;
    const int p = r >> 6, co = r & 63;
    const int t = c >> 6, j = (c >> 5) & 1, ci = c & 31;
    const int nb1 = parity ? 3 : 2;                     // pair blocks of the ky = 1 row (ky W odd <=> W odd)
    int ky, jblk, s;
    if (t < 2) { ky = 0; jblk = t; s = 0; }
    else if (t < 2 + nb1) { ky = 1; jblk = t - 2; s = parity; }
    else { ky = 2; jblk = t - 2 - nb1; s = 0; }
    const int kx = 2 * jblk + j - s - p;
    wpair[idx] = (kx >= 0 && kx <= 2) ? wt[co * 288 + (ky * 3 + kx) * 32 + ci] : __float2bfloat16(0.0f);
}

static int conv3x3_pair_taps(int w) { return (w & 1) ? 7 : 6; }

int conv3x3_pair_launch(const void* x, const void* wpair, const float* bias, void* y, int n, int h, int w, int act,
                        cudaStream_t stream) {
    const int64_t total = static_cast<int64_t>(n) * h * w;
    ISTVT_REQUIRE(x && wpair && y && n > 0 && h >= 3 && w >= 3 && (total & 1) == 0);
    ISTVT_REQUIRE(act == ISTVT_ACT_NONE || act == ISTVT_ACT_RELU);
    GemmParams p{};
    p.M = total / 2;  // pair rows on the input's grid; junk pixels dropped in the epilogue
    p.N = 128; p.K = 64;
    p.conv_w_in = w; p.conv_h_in = h; p.pair = 1;
    int t = 0;
    for (int ky = 0; ky < 3; ++ky) {
        const int64_t off = static_cast<int64_t>(ky) * w;
        const int s = static_cast<int>(off & 1);
        const int nb = s ? 3 : 2;
        for (int j = 0; j < nb; ++j) p.tap_shift[t++] = static_cast<int>((off - s) / 2) + j;
    }
    p.taps = t;
    p.b_resident = 1;   // 7 x 16 KB of weights stay in shared memory; the A ring keeps 5 of its 6 stages
    p.C = y; p.ldc = 64;
    p.bias = bias; p.residual = nullptr; p.ldr = 0;
    p.act = act; p.c_f32 = 0;
    return gemm_bf16_dispatch(x, total / 2, 64, wpair, 64 * static_cast<int64_t>(t), 64 * t, p, stream);
}

}  // namespace istvt

using namespace istvt;

extern "C" int istvt_conv3x3_pair_pack(const void* wt, void* wpair, int w_in, istvt_stream_t stream) {
    ISTVT_REQUIRE(wt && wpair && w_in >= 3);
    const int taps = conv3x3_pair_taps(w_in);
    const int total = 128 * taps * 64;
    conv3x3_pair_pack_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(wt), static_cast<__nv_bfloat16*>(wpair), w_in & 1, taps);
    count_launch();
    return launch_status();
}

extern "C" int istvt_conv3x3_pair_fwd(const void* x, const void* wpair, const float* bias, void* y, int n, int h, int w,
                                      int act, istvt_stream_t stream) {
    return conv3x3_pair_launch(x, wpair, bias, y, n, h, w, act, static_cast<cudaStream_t>(stream));
}

extern "C" int istvt_gemm_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, void* c, int64_t ldc,
                              int c_dtype, int64_t m, int n, int k, const float* bias, const float* residual,
                              int64_t ldr, int act, istvt_stream_t stream) {
    ISTVT_REQUIRE(c_dtype == ISTVT_BF16 || c_dtype == ISTVT_F32);
    ISTVT_REQUIRE(act >= ISTVT_ACT_NONE && act <= ISTVT_ACT_GELU);
    GemmParams p{};
    p.M = m; p.N = n; p.K = k;
    p.taps = 1; p.conv_w_in = 0; p.conv_h_in = 0;
    p.C = c; p.ldc = ldc;
    p.bias = bias; p.residual = residual; p.ldr = ldr;
    p.act = act; p.c_f32 = (c_dtype == ISTVT_F32);
    return gemm_bf16_dispatch(a, m, lda, w, ldw, k, p, static_cast<cudaStream_t>(stream));
}

// LayerNorm folded around a GEMM pair (see include/istvt_b200.h): both halves run on the CTA-pair kernel's TMA-store
// epilogue, so N >= 256, bf16 output, 16-byte row pitch.
extern "C" int istvt_gemm_rowstats_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, void* c, int64_t ldc,
                                       int64_t m, int n, int k, const float* bias, float* row_stats,
                                       istvt_stream_t stream) {
    ISTVT_REQUIRE(row_stats != nullptr && n >= 256 && (ldc * 2) % 16 == 0);
    ISTVT_REQUIRE((reinterpret_cast<uintptr_t>(row_stats) & 7) == 0);
    GemmParams p{};
    p.M = m; p.N = n; p.K = k;
    p.taps = 1;
    p.C = c; p.ldc = ldc;
    p.bias = bias;
    p.act = ISTVT_ACT_NONE;
    p.row_stats_out = reinterpret_cast<float2*>(row_stats);
    return gemm_bf16_dispatch(a, m, lda, w, ldw, k, p, static_cast<cudaStream_t>(stream));
}

// (mu, rstd) per row from the per-group partials of istvt_gemm_rowstats_fwd, Chan's parallel-variance combination.
__global__ void __launch_bounds__(256)
ln_stats_finalize_kernel(const float2* __restrict__ partials, float2* __restrict__ out, int64_t m, int groups, int dim,
                         float eps) {
    const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= m) return;
    const float2* st = partials + row * groups;
    float tot = 0.f;
    for (int g = 0; g < groups; ++g) tot += st[g].x;
    const float mu = tot / static_cast<float>(dim);
    float m2 = 0.f;
    for (int g = 0; g < groups; ++g) {
        const float ng = static_cast<float>(min(64, dim - 64 * g));
        const float dm = st[g].x / ng - mu;
        m2 += st[g].y + ng * dm * dm;
    }
    out[row] = make_float2(mu, rsqrtf(m2 / static_cast<float>(dim) + eps));
}

extern "C" int istvt_ln_stats_finalize(const float* partials, float* mu_rstd, int64_t m, int dim, float eps,
                                       istvt_stream_t stream) {
    ISTVT_REQUIRE(partials && mu_rstd && m > 0 && dim > 0);
    ISTVT_REQUIRE(((reinterpret_cast<uintptr_t>(partials) | reinterpret_cast<uintptr_t>(mu_rstd)) & 7) == 0);
    ln_stats_finalize_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(partials), reinterpret_cast<float2*>(mu_rstd), m, (dim + 63) / 64, dim, eps);
    count_launch();
    return launch_status();
}

extern "C" int istvt_gemm_lnfold_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, void* c, int64_t ldc,
                                     int64_t m, int n, int k, const float* mu_rstd, const float* w_rowsum,
                                     const float* shift, istvt_stream_t stream) {
    ISTVT_REQUIRE(mu_rstd != nullptr && w_rowsum != nullptr && shift != nullptr && n >= 256 && (ldc * 2) % 16 == 0);
    ISTVT_REQUIRE(((reinterpret_cast<uintptr_t>(mu_rstd) & 7) | (reinterpret_cast<uintptr_t>(w_rowsum) & 15) |
                   (reinterpret_cast<uintptr_t>(shift) & 15)) == 0);
    GemmParams p{};
    p.M = m; p.N = n; p.K = k;
    p.taps = 1;
    p.C = c; p.ldc = ldc;
    p.act = ISTVT_ACT_NONE;
    p.ln_stats = reinterpret_cast<const float2*>(mu_rstd); p.ln_c = w_rowsum; p.ln_d = shift;
    return gemm_bf16_dispatch(a, m, lda, w, ldw, k, p, static_cast<cudaStream_t>(stream));
}

// MLP fusions of the training step (see include/istvt_b200.h)
extern "C" int istvt_gemm_act_dual_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, void* c_act, int64_t ldc,
                                       void* c_pre, int64_t ldc_pre, int64_t m, int n, int k, const float* bias, int act,
                                       istvt_stream_t stream) {
    ISTVT_REQUIRE(c_pre != nullptr && n >= 256 && (ldc * 2) % 16 == 0 && (ldc_pre * 2) % 16 == 0);
    ISTVT_REQUIRE((reinterpret_cast<uintptr_t>(c_pre) & 15) == 0);
    ISTVT_REQUIRE(act >= ISTVT_ACT_NONE && act <= ISTVT_ACT_GELU);
    GemmParams p{};
    p.M = m; p.N = n; p.K = k;
    p.taps = 1;
    p.C = c_act; p.ldc = ldc;
    p.C2 = c_pre; p.ldc2 = ldc_pre;
    p.bias = bias;
    p.act = act;
    return gemm_bf16_dispatch(a, m, lda, w, ldw, k, p, static_cast<cudaStream_t>(stream));
}

extern "C" int istvt_gemm_dgelu_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, void* c, int64_t ldc,
                                    const void* pre, int64_t ld_pre, int64_t m, int n, int k, istvt_stream_t stream) {
    ISTVT_REQUIRE(pre != nullptr && n >= 256 && (ldc * 2) % 16 == 0 && ld_pre % 8 == 0 && ld_pre >= n);
    ISTVT_REQUIRE((reinterpret_cast<uintptr_t>(pre) & 15) == 0);
    GemmParams p{};
    p.M = m; p.N = n; p.K = k;
    p.taps = 1;
    p.C = c; p.ldc = ldc;
    p.mul_pre = pre; p.ld_pre = ld_pre;
    p.act = ISTVT_ACT_NONE;
    return gemm_bf16_dispatch(a, m, lda, w, ldw, k, p, static_cast<cudaStream_t>(stream));
}

// split-K plan shared by the two weight-gradient entry points
static void plan_splitk(GemmParams& p, int64_t m, int n) {
    const int num_kb = (p.K + 63) / 64;
    const int64_t mn_tiles = ((m + 255) / 256) * ((n + 255) / 256);
    int64_t want = (4 * (sm_count() / 2) + mn_tiles - 1) / mn_tiles;   // ~4 waves of (tile, range) items
    if (want < 2) want = 2;
    int kb_per = static_cast<int>((num_kb + want - 1) / want);
    if (kb_per < 8) kb_per = 8;
    if (kb_per > num_kb) kb_per = num_kb;
    p.kb_per_split = kb_per;
    p.split_k = (num_kb + kb_per - 1) / kb_per;
}

// Weight gradient without transposed copies: dW[n_out, k_in] += dY[rows, n_out]^T X[rows, k_in], both operands read
// in place as MN-major tiles (the token rows are the contraction dimension).  Always split-K + red.global.add.
extern "C" int istvt_gemm_wgrad_accum(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, float* dw,
                                      int64_t ld_dw, int64_t rows, int n_out, int k_in, istvt_stream_t stream) {
    ISTVT_REQUIRE(dy && x && dw);
    ISTVT_REQUIRE(rows > 0 && rows < (int64_t(1) << 31) && n_out > 0 && k_in > 0);
    ISTVT_REQUIRE(n_out % 8 == 0 && k_in % 8 == 0 && ld_dy % 8 == 0 && ld_x % 8 == 0 && ld_dw % 4 == 0);
    ISTVT_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dw) & 15) == 0);
    GemmParams p{};
    p.M = n_out; p.N = k_in; p.K = static_cast<int>(rows);
    p.taps = 1;
    p.C = dw; p.ldc = ld_dw;
    p.act = ISTVT_ACT_NONE; p.c_f32 = 1;
    p.mn_major = 1;
    const int num_kb = (p.K + 63) / 64;
    if (num_kb < 2) return ISTVT_ERR_UNSUPPORTED;      // the accumulate epilogue needs >= 2 non-empty K ranges
    plan_splitk(p, n_out, k_in);
    if (p.split_k < 2) {
        p.kb_per_split = (num_kb + 1) / 2;
        p.split_k = (num_kb + p.kb_per_split - 1) / p.kb_per_split;   // == 2
    }
    return launch_gemm_2cta_mn(dy, ld_dy, x, ld_x, p, static_cast<cudaStream_t>(stream));
}

// Weight-gradient GEMM: C[m, n] += sum_k A[m, k] * W[n, k] with a very long K (the token rows) and a small
// output: split-K over CTA pairs, partial products accumulated with red.global.add into the fp32 C.
extern "C" int istvt_gemm_splitk_accum(const void* a, int64_t lda, const void* w, int64_t ldw, float* c, int64_t ldc,
                                       int64_t m, int n, int64_t k, istvt_stream_t stream) {
    ISTVT_REQUIRE(a && w && c);
    ISTVT_REQUIRE(m > 0 && n > 0 && k > 0 && k < (int64_t(1) << 31));
    ISTVT_REQUIRE(n % 4 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 4 == 0);
    ISTVT_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(c) & 15) == 0);
    GemmParams p{};
    p.M = m; p.N = n; p.K = static_cast<int>(k);
    p.taps = 1;
    p.C = c; p.ldc = ldc;
    p.bias = nullptr; p.residual = nullptr; p.ldr = 0;
    p.act = ISTVT_ACT_NONE; p.c_f32 = 1;
    const int num_kb = (p.K + 63) / 64;
    const int64_t mn_tiles = ((m + 255) / 256) * ((n + 255) / 256);
    // enough (tile, range) items for ~4 waves over the CTA pairs, at least 8 k-blocks per range
    int64_t want = (4 * (sm_count() / 2) + mn_tiles - 1) / mn_tiles;
    if (want < 2) want = 2;
    int kb_per = static_cast<int>((num_kb + want - 1) / want);
    if (kb_per < 8) kb_per = 8;
    if (kb_per > num_kb) kb_per = num_kb;
    p.kb_per_split = kb_per;
    p.split_k = (num_kb + kb_per - 1) / kb_per;
    if (p.split_k < 2) { p.split_k = 2; p.kb_per_split = (num_kb + 1) / 2; if (p.kb_per_split * 1 >= num_kb) { p.split_k = 1; } }
    if (p.split_k == 1) {
        // degenerate (K <= 64): plain accumulate through the residual epilogue
        p.split_k = 0; p.kb_per_split = 0; p.residual = c; p.ldr = ldc;
    }
    return launch_gemm_2cta(a, lda, w, ldw, p, static_cast<cudaStream_t>(stream));
}
