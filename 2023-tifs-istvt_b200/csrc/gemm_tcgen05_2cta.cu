// CTA-pair tcgen05 GEMM for sm_100a (cta_group::2):  C = act(A · Wᵀ + bias) + residual,  N >= 256.
//
// Why pairs: profiles/r1a showed the 1-CTA 128x256 tile operand-feed bound (L2->SM 11.8 TB/s, tensor pipe
// 49 %): it pulls (128 + 256) operand rows per 128x256 outputs.  A cluster of two CTAs computes a 256x256 tile
// with ONE tcgen05.mma.cta_group::2 stream: each CTA stages its own 128 A rows and only HALF of the B tile
// (128 W rows); the tensor cores of both SMs read both halves.  Operand traffic per output drops by 1/3 and
// the smem ring gets 5 stages of 32 KB instead of 4 of 48 KB.
//
// Roles per CTA (640 threads): warp 0 = TMA producer (both CTAs; transaction bytes are credited to the
// leader's full barrier), warp 1 = MMA issuer (LEADER CTA only, one lane), warp 2 = TMEM allocator,
// warps 4..19 = epilogue of this CTA's 128 accumulator rows (32 rows x 64 columns each).  Barriers live at identical smem offsets in both
// CTAs: full[s] (leader's copy used), empty[s] / tmem_full[a] (tcgen05.commit multicast to both CTAs),
// tmem_empty[a] (leader's copy, 2 x 8 epilogue-warp arrivals, the peer arrives remotely via mapa).
#include "gemm_common.cuh"

#include <stdlib.h>

namespace istvt {

// Debug timeline (compiled only with -DISTVT_GEMM_TRACE, see tools/gemm_trace.py): cluster 0 records globaltimer
// stamps of its first tiles — [tile][0] producer: first TMA of the tile issued, [1] last TMA issued, [2] issuer: accumulator
// free, [3] first k-block landed, [4] last MMA + commits issued, [5] epilogue warp 4: tmem_full seen, [6] its TMEM reads
// done (arrive on tmem_empty), [7] its last store issued.
#ifdef ISTVT_GEMM_TRACE
__device__ unsigned long long* g_gemm_trace = nullptr;
constexpr int TRACE_TILES = 24;
__device__ __forceinline__ void trace_stamp(bool on, int tile_it, int slot) {
    if (on && tile_it < TRACE_TILES && g_gemm_trace != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_gemm_trace[tile_it * 8 + slot] = t;
    }
}
#define ISTVT_TRACE(on, it, slot) trace_stamp(on, it, slot)
#else
#define ISTVT_TRACE(on, it, slot) ((void)0)
#endif

constexpr int G2_BN = 256;                    // cluster tile N (UMMA N), 128 W rows staged per CTA
constexpr int G2_BK = 64;                     // 128-byte swizzle atom
constexpr int G2_A_BYTES = GEMM_BLOCK_M * G2_BK * 2;   // 16 KB
constexpr int G2_B_BYTES = (G2_BN / 2) * G2_BK * 2;    // 16 KB
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
// epilogue modes: bf16 output without residual / generic (fp32 output, residual through registers, split-K) /
// in-place fp32 residual update by TMA reduce-add
// in-place fp32 residual update by TMA reduce-add / bf16 output whose TMA-store boxes are 64 columns (128 bytes) wide
// EPI_DUAL: bf16 output twice (pre-activation and activated), two 32-column boxes per half from a 4 KB slab
// EPI_DGELU: EPI_PLAIN128 whose accumulator is multiplied by gelu'(mul_pre) first
constexpr int EPI_PLAIN = 0, EPI_GENERIC = 1, EPI_REDUCE = 2, EPI_PLAIN128 = 3, EPI_DUAL = 4, EPI_DGELU = 5;
template <int EW, int EPI> struct G2Cfg {
    static constexpr int THREADS = 128 + EW * 32;
    static constexpr int SLAB = EPI == EPI_PLAIN ? EPI_SLAB_PLAIN_BYTES : EPI_SLAB_BYTES;   // PLAIN128: 32 rows x 128 B
    // 6 ring stages whenever they fit: 16 warps with the 2 KB bf16 slabs, or 8 warps with the 4 KB fp32 slabs
    static constexpr int STAGES = (EW == 16 && EPI != EPI_PLAIN) ? 5 : 6;
    static constexpr int PASSES = (G2_BN / EPI_COLS) / (EW / 4);   // 64-column passes per epilogue warp per tile
    static constexpr int SMEM_BYTES = STAGES * G2_STAGE_BYTES + EW * SLAB + 1024 + 256;
};
constexpr int G2_TMEM_COLS = 512;             // 2 accumulator buffers x 256 fp32 columns

// Static tile schedule of one cluster, tile = ks * (m_tiles * n_tiles) + m_blk * n_tiles + n_blk, advanced by the
// cluster count with 32-bit adds.  The first version recomputed (ks, m_blk, n_blk) from the 64-bit tile index with two
// 64-bit divisions per tile in every role: on the single MMA-issuing thread that was 0.57 us between the last MMA of
// one tile and the first of the next — 9 % of a K = 728 tile (tools/gemm_trace.py, profiles/README.md r3u).
struct TileIter {
    int ks, m_blk, n_blk;          // current tile
    int s_ks, s_m, s_n;            // decomposition of the step (number of clusters)
    int m_tiles, n_tiles;
    int64_t tile, total, step;
    __device__ __forceinline__ TileIter(int64_t start, int64_t step_, int m_tiles_, int n_tiles_, int64_t total_)
        : m_tiles(m_tiles_), n_tiles(n_tiles_), tile(start), total(total_), step(step_) {
        const int64_t mn = static_cast<int64_t>(m_tiles) * n_tiles;
        ks = static_cast<int>(start / mn);
        const int64_t r = start - ks * mn;
        m_blk = static_cast<int>(r / n_tiles);
        n_blk = static_cast<int>(r - static_cast<int64_t>(m_blk) * n_tiles);
        s_ks = static_cast<int>(step / mn);
        const int64_t sr = step - s_ks * mn;
        s_m = static_cast<int>(sr / n_tiles);
        s_n = static_cast<int>(sr - static_cast<int64_t>(s_m) * n_tiles);
    }
    __device__ __forceinline__ bool valid() const { return tile < total; }
    __device__ __forceinline__ void next() {
        tile += step;
        n_blk += s_n;
        if (n_blk >= n_tiles) { n_blk -= n_tiles; ++m_blk; }
        m_blk += s_m;
        if (m_blk >= m_tiles) { m_blk -= m_tiles; ++ks; }
        ks += s_ks;
    }
};

template <int EW, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2Cfg<EW, EPI>::THREADS, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                         const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_c2,
                         const GemmParams p) {
    constexpr bool PLAIN_BF16 = EPI == EPI_PLAIN || EPI == EPI_PLAIN128 || EPI == EPI_DUAL || EPI == EPI_DGELU;
    using Cfg = G2Cfg<EW, EPI>;
    constexpr int G2_STAGES = Cfg::STAGES;
    constexpr int G2_EPI_WARPS = EW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + G2_STAGES * G2_A_BYTES;
    uint8_t* smem_epi = smem + G2_STAGES * G2_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + G2_EPI_WARPS * Cfg::SLAB);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + G2_STAGES;
    uint64_t* tmem_full = bars + 2 * G2_STAGES;
    uint64_t* tmem_empty = bars + 2 * G2_STAGES + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * G2_STAGES + 4);

    const int lane = threadIdx.x & 31;
    // (Placing the single-thread roles in the four HIGHEST warp ids — the issue arbiter is said to prefer the highest
    //  eligible warp id — measured neutral: layer GEMM sum 2.322-2.326 vs 2.315-2.319 ms, profiles/README.md r6o.)
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();          // 0 = leader
    const uint32_t cluster = cluster_id_x();
    const uint32_t n_clusters = num_clusters_x();

    const int num_kb = (p.K + G2_BK - 1) / G2_BK;
    const int n_tiles = (p.N + G2_BN - 1) / G2_BN;
    const int m_tiles = static_cast<int>((p.M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M));
    const int splits = p.split_k > 1 ? p.split_k : 1;
    const int kb_per = p.split_k > 1 ? p.kb_per_split : num_kb;
    const int64_t mn_tiles = static_cast<int64_t>(m_tiles) * n_tiles;
    const int64_t total_tiles = mn_tiles * splits;     // tile = split * mn_tiles + (m_blk * n_tiles + n_blk)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < G2_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);     // leader's arrive.expect_tx; bytes of both CTAs' loads
            mbar_init(&empty_bar[s], 1);    // one multicast commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);                    // one multicast commit
            mbar_init(&tmem_empty[s], 2 * G2_EPI_WARPS);    // epilogue warps of both CTAs
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_2cta(tmem_holder, G2_TMEM_COLS);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    __syncthreads();       // CTA-level ordering of the tmem_holder write (compute-sanitizer racecheck does not model
                           // barrier.cluster as a shared-memory barrier; once per kernel)
    cluster_sync_all();    // barrier inits + TMEM allocation visible in both CTAs before any remote arrive / TMA
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    // columns of the last N tile rounded up to 16; each CTA stages half of them
    auto tile_n_eff = [&](int n_blk) {
        int n_eff = p.N - n_blk * G2_BN;
        return n_eff >= G2_BN ? G2_BN : ((n_eff + 15) & ~15);
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs, one elected thread each) =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            [[maybe_unused]] int trace_it = 0;
            for (TileIter ti(cluster, n_clusters, m_tiles, n_tiles, total_tiles); ti.valid(); ti.next(), ++trace_it) {
                const int ks = ti.ks;
                const int n_blk = ti.n_blk;
                const int m0 = ti.m_blk * (2 * GEMM_BLOCK_M) + static_cast<int>(rank) * GEMM_BLOCK_M;
                const int n0 = n_blk * G2_BN + static_cast<int>(rank) * (tile_n_eff(n_blk) >> 1);
                const int kb_end = (ks + 1) * kb_per < num_kb ? (ks + 1) * kb_per : num_kb;
                for (int kb = ks * kb_per; kb < kb_end; ++kb) {
                    mbar_wait_hot(&empty_bar[stage], phase ^ 1);
                    if (kb == ks * kb_per) ISTVT_TRACE(cluster == 0 && rank == 0, trace_it, 0);
                    if (kb == kb_end - 1) ISTVT_TRACE(cluster == 0 && rank == 0, trace_it, 1);
#ifdef ISTVT_GEMM_TRACE
                    // experiment (results are garbage): after the ring's first fill, signal the slots full WITHOUT loading,
                    // so that the MMAs re-read stale shared memory with no concurrent TMA fills (ISTVT_TRACE_NOTMA=1)
                    if (p.trace_no_tma == 1 && (trace_it > 0 || kb >= G2_STAGES)) {
                        if (rank == 0) mbar_arrive(&full_bar[stage]);
                        if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
#endif
#ifdef ISTVT_GEMM_TRACE
                    // ISTVT_TRACE_NOTMA=2 / 3: after the first ring fill only the A / only the B boxes are fetched (the
                    // other operand is re-read stale): how much of the k-block time is the fill of each operand
                    if (p.trace_no_tma >= 2 && !p.mn_major && (trace_it > 0 || kb >= G2_STAGES)) {
                        const bool only_a = p.trace_no_tma == 2;
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (only_a ? G2_A_BYTES : G2_B_BYTES));
                        if (only_a) tma_load_2d_2cta(smem_a + stage * G2_A_BYTES, &tm_a, &full_bar[stage], kb * G2_BK, m0);
                        else        tma_load_2d_2cta(smem_b + stage * G2_B_BYTES, &tm_b, &full_bar[stage], kb * G2_BK, n0);
                        if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
#endif
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
                    if (!p.mn_major) {
                        tma_load_2d_2cta(smem_a + stage * G2_A_BYTES, &tm_a, &full_bar[stage], kb * G2_BK, m0);
                        tma_load_2d_2cta(smem_b + stage * G2_B_BYTES, &tm_b, &full_bar[stage], kb * G2_BK, n0);
                    } else {
                        // MN-major: a 128 (mn) x 64 (k) operand tile = two boxes of 64 k-rows x 64 contiguous mn
                        // elements (128 B, SW128), 8 KB each, placed back to back (LBO = 8192)
                        tma_load_2d_2cta(smem_a + stage * G2_A_BYTES, &tm_a, &full_bar[stage], m0, kb * G2_BK);
                        tma_load_2d_2cta(smem_a + stage * G2_A_BYTES + 8192, &tm_a, &full_bar[stage], m0 + 64, kb * G2_BK);
                        tma_load_2d_2cta(smem_b + stage * G2_B_BYTES, &tm_b, &full_bar[stage], n0, kb * G2_BK);
                        tma_load_2d_2cta(smem_b + stage * G2_B_BYTES + 8192, &tm_b, &full_bar[stage], n0 + 64, kb * G2_BK);
                    }
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (leader CTA, ONE elected thread) =====================
        // The issue loop is the critical path of the kernel (profiles/r1f: ~26 SASS instructions per tcgen05.mma
        // and ~75 per k-block made the issuing thread slower than the tensor core), so it is kept minimal:
        // descriptors are a precomputed constant plus the stage's address field, the ragged last k-block is
        // peeled, and the only per-k-block extras are one barrier wait and one commit.
        if (elect_one()) {
            // K-major: SBO = 1024 (8 rows x 128 B), k-step = +32 B.  MN-major: LBO = 8192 (next 64-wide mn atom),
            // SBO = 1024 (next 8 k-rows), k-step = 16 k-rows = +2048 B.
            const uint64_t desc_hi = p.mn_major ? make_smem_desc(0, 8192, 1024, SWZ_128B) : make_smem_desc(0, 0, 1024, SWZ_128B);
            const uint64_t kadv = p.mn_major ? (2048u >> 4) : (32u >> 4);
            const uint32_t a_field0 = (smem_u32(smem_a) & 0x3FFFFu) >> 4;
            const uint32_t b_field0 = (smem_u32(smem_b) & 0x3FFFFu) >> 4;
            int last_steps = (p.K - (num_kb - 1) * G2_BK + UMMA_K - 1) / UMMA_K;   // 1..4 MMAs in the last k-block
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            [[maybe_unused]] int trace_it = 0;
            for (TileIter ti(cluster, n_clusters, m_tiles, n_tiles, total_tiles); ti.valid(); ti.next(), ++trace_it) {
                const int ks = ti.ks;
                const int n_blk = ti.n_blk;
                const uint32_t mnm = p.mn_major ? 1u : 0u;
                const uint32_t idesc = make_idesc_bf16(2 * GEMM_BLOCK_M, static_cast<uint32_t>(tile_n_eff(n_blk)), mnm, mnm);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                ISTVT_TRACE(cluster == 0, trace_it, 2);
                const uint32_t d_tmem = tmem_base + acc * G2_BN;
                const int kb_begin = ks * kb_per;
                const int kb_end = (ks + 1) * kb_per < num_kb ? (ks + 1) * kb_per : num_kb;
                const int tail_steps = kb_end == num_kb ? last_steps : G2_BK / UMMA_K;
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait_hot(&full_bar[stage], phase);
                    tc_fence_after();
                    if (kb == kb_begin) ISTVT_TRACE(cluster == 0, trace_it, 3);
                    const uint64_t a_desc = desc_hi | (a_field0 + stage * (G2_A_BYTES >> 4));
                    const uint64_t b_desc = desc_hi | (b_field0 + stage * (G2_B_BYTES >> 4));
                    if (kb != kb_end - 1) {
                        umma_f16_ss_2cta(d_tmem, a_desc, b_desc, idesc, kb != kb_begin ? 1u : 0u);
                        umma_f16_ss_2cta(d_tmem, a_desc + kadv, b_desc + kadv, idesc, 1u);
                        umma_f16_ss_2cta(d_tmem, a_desc + 2 * kadv, b_desc + 2 * kadv, idesc, 1u);
                        umma_f16_ss_2cta(d_tmem, a_desc + 3 * kadv, b_desc + 3 * kadv, idesc, 1u);
                        umma_commit_2cta(&empty_bar[stage], 3);                  // free the slot in both CTAs
                    } else {
                        for (int k = 0; k < tail_steps; ++k)
                            umma_f16_ss_2cta(d_tmem, a_desc + k * kadv, b_desc + k * kadv, idesc,
                                             (kb != kb_begin || k != 0) ? 1u : 0u);
                        umma_commit_2cta(&empty_bar[stage], 3);
                        umma_commit_2cta(&tmem_full[acc], 3);                    // accumulator ready in both CTAs
                        ISTVT_TRACE(cluster == 0, trace_it, 4);
                    }
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue (this CTA's 128 rows x 256 columns) =====================
        const int ew = warp - 4;
        const int quad = warp & 3;          // TMEM lane quadrant this warp may access
        constexpr int PASSES = Cfg::PASSES;
        const int colw = (ew >> 2) * (EPI_COLS * PASSES);   // first tile column of this warp
        const uint32_t slab = smem_u32(smem_epi + ew * Cfg::SLAB);
        int acc = 0;
        uint32_t acc_phase = 0;
        [[maybe_unused]] int trace_it = 0;
        [[maybe_unused]] const bool tracer = cluster == 0 && rank == 0 && ew == 0 && lane == 0;
        for (TileIter ti(cluster, n_clusters, m_tiles, n_tiles, total_tiles); ti.valid(); ti.next(), ++trace_it) {
            const int m_blk = ti.m_blk;
            const int n_blk = ti.n_blk;
            const int64_t m = static_cast<int64_t>(m_blk) * (2 * GEMM_BLOCK_M) + rank * GEMM_BLOCK_M + quad * 32 + lane;
            int drow_t[8];
            const int drow_lane = m < p.M ? static_cast<int>(m) : -1;
            epilogue_rows(drow_lane, lane, drow_t);
            if constexpr (EPI == EPI_GENERIC) epilogue_prefetch_residual(p, drow_lane, n_blk * G2_BN + colw, EPI_COLS * PASSES);
            if (EPI == EPI_DGELU && drow_lane >= 0) {
                // ff2 data gradient: this lane's 64 pre-activations of the tile (128 bytes) into L2 before the wait
                const int nf = n_blk * G2_BN + colw;
                for (int n = nf; n < nf + EPI_COLS * PASSES && n < p.N; n += 32)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const __nv_bfloat16*>(p.mul_pre) +
                                                                  static_cast<int64_t>(drow_lane) * p.ld_pre + n));
            }
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            ISTVT_TRACE(tracer, trace_it, 5);
            uint64_t* te = &tmem_empty[acc];
#ifdef ISTVT_GEMM_TRACE
            if (p.trace_no_epi == 1) {     // timing experiment: no TMEM reads, no math, no stores
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(te, 0);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
#endif
#pragma unroll 1
            for (int ps = 0; ps < PASSES; ++ps) {
                const int col0 = colw + ps * EPI_COLS;
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * G2_BN + col0;
                auto release = [&]() {
                    if (ps != PASSES - 1) return;
                    // all TMEM reads of this warp for this accumulator are done: tell the leader's MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(te, 0);
                    ISTVT_TRACE(tracer, trace_it, 6);
                };
                if constexpr (EPI == EPI_REDUCE)
                    gemm_epilogue_reduce_64(p, &tm_c, taddr, slab,
                                            static_cast<int>(m_blk * (2 * GEMM_BLOCK_M) + rank * GEMM_BLOCK_M + quad * 32),
                                            n_blk * G2_BN + col0, lane, release);
                else if (PLAIN_BF16 && p.epi_tma)
                    gemm_epilogue_tma_bf16_64<(EPI == EPI_PLAIN128 || EPI == EPI_DGELU) ? 64 : 32, EPI == EPI_DUAL, EPI == EPI_DGELU>(
                        p, &tm_c, &tm_c2, taddr, slab,
                        static_cast<int>(m_blk * (2 * GEMM_BLOCK_M) + rank * GEMM_BLOCK_M + quad * 32),
                        n_blk * G2_BN + col0, lane, release);
                else
                    gemm_epilogue_64<PLAIN_BF16>(p, taddr, slab, drow_lane, drow_t, n_blk * G2_BN + col0, lane, release);
            }
            ISTVT_TRACE(tracer, trace_it, 7);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    if (EPI == EPI_REDUCE || (PLAIN_BF16 && p.epi_tma)) {
        if (warp >= 4 && lane == 0) tma_store_wait0();     // TMA stores / reduce-adds issued by this lane have completed
    }
    // No CTA may exit (or free TMEM) while its peer can still multicast-arrive on its barriers or read its smem.
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, G2_TMEM_COLS);
    }
}

static int launch_gemm_2cta_maps(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const GemmParams& p, cudaStream_t stream);

int launch_gemm_2cta_mn(const void* a, int64_t lda, const void* w, int64_t ldw, const GemmParams& p, cudaStream_t stream) {
    CUtensorMap tm_a, tm_b;
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.M), static_cast<uint64_t>(p.K)};
        const uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
        const uint32_t box[2] = {64, G2_BK};
        int rc = encode_tmap(&tm_a, a, ISTVT_BF16, 2, dims, strides, box, 3);
        if (rc != ISTVT_OK) return rc;
    }
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(p.K)};
        const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
        const uint32_t box[2] = {64, G2_BK};
        int rc = encode_tmap(&tm_b, w, ISTVT_BF16, 2, dims, strides, box, 3);
        if (rc != ISTVT_OK) return rc;
    }
    return launch_gemm_2cta_maps(tm_a, tm_b, p, stream);
}

int launch_gemm_2cta(const void* a, int64_t lda, const void* w, int64_t ldw, const GemmParams& p, cudaStream_t stream) {
    CUtensorMap tm_a, tm_b;
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.M)};
        const uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
        const uint32_t box[2] = {G2_BK, GEMM_BLOCK_M};
        int rc = encode_tmap(&tm_a, a, ISTVT_BF16, 2, dims, strides, box, 3);
        if (rc != ISTVT_OK) return rc;
    }
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.N)};
        const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
        const uint32_t box[2] = {G2_BK, G2_BN / 2};
        int rc = encode_tmap(&tm_b, w, ISTVT_BF16, 2, dims, strides, box, 3);
        if (rc != ISTVT_OK) return rc;
    }
    return launch_gemm_2cta_maps(tm_a, tm_b, p, stream);
}

static int launch_gemm_2cta_maps(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const GemmParams& p_in, cudaStream_t stream) {
    const GemmParams& p0 = p_in;
    // (A TMA L2 prefetch of the next m-block's A rows one tile ahead was measured SLOWER — layer GEMM sum 2.42 ->
    //  2.57 ms: the kernel is bound by L2->SM delivery, not by first-touch DRAM latency, and the prefetch requests
    //  compete for the same L2 slices.  profiles/README.md r3a.  A second attempt on the LSU path — an otherwise idle
    //  warp issuing prefetch.global.L2 for the m-block two tiles ahead — was worse still: 2.48 -> 3.08 ms, r3g.)
    const int n_tiles = (p0.N + G2_BN - 1) / G2_BN;
    const int64_t m_tiles = (p0.M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);
    const int64_t total = m_tiles * n_tiles * (p0.split_k > 1 ? p0.split_k : 1);
    int64_t clusters = sm_count() / 2;
    if (total < clusters) clusters = total;
    // Epilogue warps per CTA: 16 (32 rows x 64 columns each) for the bf16 path — halves the per-tile epilogue
    // latency, which matters for the K = 512 / 728 GEMMs — and 8 for the fp32 / residual path, whose register
    // footprint does not fit 640 threads (measured: profiles/README.md r1i).  ISTVT_G2_EPI_WARPS overrides.
    static const int ew_env = []() {
        const char* e = getenv("ISTVT_G2_EPI_WARPS");
        return e ? atoi(e) : 0;
    }();
    // ISTVT_G2_TMASTORE=0: bf16 epilogue with per-lane LDS + STG instead of TMA box stores (A/B measurements)
    static const int tma_env = []() { const char* e = getenv("ISTVT_G2_TMASTORE"); return e ? atoi(e) : 1; }();
    GemmParams p = p_in;
#ifdef ISTVT_GEMM_TRACE
    {
        const char* e = getenv("ISTVT_TRACE_NOTMA");
        p.trace_no_tma = e ? atoi(e) : 0;
        const char* f = getenv("ISTVT_TRACE_NOEPI");      // 1: the epilogue warps only release the accumulator
        p.trace_no_epi = f ? atoi(f) : 0;
    }
#endif
    const bool plain = !p.c_f32 && p.residual == nullptr;
    // In-place fp32 residual update (inference s_out / ff2): the addition is done by the L2 through a TMA reduce-add.
    // ISTVT_G2_REDUCE=0 keeps the register-path epilogue (A/B measurements).
    static const bool reduce_env = []() { const char* e = getenv("ISTVT_G2_REDUCE"); return !e || atoi(e) != 0; }();
    const bool reduce = reduce_env && p.c_f32 && p.residual != nullptr && p.residual == p.C && p.ldr == p.ldc &&
                        p.split_k <= 1 && !p.mn_major && (p.ldc * 4) % 16 == 0;
    CUtensorMap tm_c = tm_a;     // placeholder for the modes that do not use it
    if (reduce) {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(p.M)};
        const uint64_t strides[1] = {static_cast<uint64_t>(p.ldc) * 4};
        const uint32_t box[2] = {32, 32};      // 32 fp32 columns (128 B, SW128) x the warp's 32 rows
        int rc = encode_tmap(&tm_c, p.C, ISTVT_F32, 2, dims, strides, box, 3);
        if (rc != ISTVT_OK) return rc;
    }
    const int ew = reduce ? 16 : (ew_env == 8 || ew_env == 16 ? ew_env : (plain ? 16 : 8));
    p.epi_tma = plain && tma_env && !p.mn_major && p.split_k <= 1 && (p.ldc * 2) % 16 == 0;
    if ((p.row_stats_out != nullptr || p.ln_stats != nullptr) && !p.epi_tma) return ISTVT_ERR_UNSUPPORTED;
    // ISTVT_G2_STORE128=0: 32-column store boxes (2 KB slabs, six ring stages) instead of 64-column ones (4 KB slabs,
    // five stages) — A/B measurements
    static const bool store128_env = []() { const char* e = getenv("ISTVT_G2_STORE128"); return !e || atoi(e) != 0; }();
    const bool dual = p.C2 != nullptr;
    if (dual && !(p.epi_tma && (p.ldc2 * 2) % 16 == 0)) return ISTVT_ERR_UNSUPPORTED;
    const bool dgelu = p.mul_pre != nullptr;
    if (dgelu && (!p.epi_tma || dual)) return ISTVT_ERR_UNSUPPORTED;
    const bool store128 = p.epi_tma && (store128_env || dgelu) && !dual;
    CUtensorMap tm_c2 = tm_a;    // placeholder unless dual
    if (dual) {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(p.M)};
        const uint64_t strides[1] = {static_cast<uint64_t>(p.ldc2) * 2};
        const uint32_t box[2] = {32, 32};
        int rc = encode_tmap(&tm_c2, p.C2, ISTVT_BF16, 2, dims, strides, box, 2);
        if (rc != ISTVT_OK) return rc;
    }
    if (p.epi_tma) {
        const uint64_t dims[2] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(p.M)};
        const uint64_t strides[1] = {static_cast<uint64_t>(p.ldc) * 2};
        // the warp's 32 rows x 32 bf16 columns (64 B, SW64) or x 64 columns (128 B, SW128)
        const uint32_t box[2] = {store128 ? 64u : 32u, 32};
        int rc = encode_tmap(&tm_c, p.C, ISTVT_BF16, 2, dims, strides, box, store128 ? 3 : 2);
        if (rc != ISTVT_OK) return rc;
    }
    const unsigned grid = static_cast<unsigned>(2 * clusters);
#define ISTVT_G2_LAUNCH(EWV, EPIV)                                                                                 \
    do {                                                                                                           \
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<EWV, EPIV>,                                 \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, G2Cfg<EWV, EPIV>::SMEM_BYTES)); \
        gemm_tcgen05_2cta_kernel<EWV, EPIV><<<grid, G2Cfg<EWV, EPIV>::THREADS, G2Cfg<EWV, EPIV>::SMEM_BYTES, stream>>>(   \
            tm_a, tm_b, tm_c, tm_c2, p);                                                                           \
    } while (0)
    if (reduce) {
        ISTVT_G2_LAUNCH(16, EPI_REDUCE);
    } else if (dual) {
        ISTVT_G2_LAUNCH(16, EPI_DUAL);
    } else if (dgelu) {
        ISTVT_G2_LAUNCH(16, EPI_DGELU);
    } else if (ew == 16) {
        if (store128) ISTVT_G2_LAUNCH(16, EPI_PLAIN128);
        else if (plain) ISTVT_G2_LAUNCH(16, EPI_PLAIN);
        else ISTVT_G2_LAUNCH(16, EPI_GENERIC);
    } else {
        if (store128) ISTVT_G2_LAUNCH(8, EPI_PLAIN128);
        else if (plain) ISTVT_G2_LAUNCH(8, EPI_PLAIN);
        else ISTVT_G2_LAUNCH(8, EPI_GENERIC);
    }
#undef ISTVT_G2_LAUNCH
    count_launch();
    return launch_status();
}

}  // namespace istvt

#ifdef ISTVT_GEMM_TRACE
extern "C" int istvt_debug_gemm_trace(unsigned long long* buf) {   // debug builds only; not part of the ABI header
    return static_cast<int>(cudaMemcpyToSymbol(istvt::g_gemm_trace, &buf, sizeof(buf)));
}
#endif
