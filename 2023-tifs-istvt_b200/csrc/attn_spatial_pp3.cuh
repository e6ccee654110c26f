// Spatial attention, inference kernel with THREE query tiles in flight per CTA (included by attn_spatial.cu, after
// attn_spatial_pp.cuh whose softmax helpers it shares).  network/vivit/module.py:84-91.
// OPT-IN (ISTVT_SA_KERNEL=pp3): parity-green, but it measures the same as the two-tile kernel (0.258 vs 0.260 ms per
// launch at the C2 size, XU pipe 53 % in both, profiles/README.md r7r / r7u) — a third softmax warp per scheduler does
// not raise the MUFU occupancy; the warps are not parked on barriers (pv_done 7 %, tcgen05.wait::st 6 %, s_full 4 % of
// their samples) but issue ~6 instructions per score at 50 % issue utilisation with `wait` as the top stall.
//
// Why: with two tiles in flight (attn_spatial_pp_kernel) a scheduler holds two softmax warps; whenever one of them is in
// its load / max / hand-off phase the other has the MUFU pipe to itself but cannot fill it alone — the XU pipe is 54 %
// busy, `wait` (fixed-latency dependencies) is the top stall (profiles/r7m_attn_spatial_pp_ncu_source.txt).  A frame has
// 362 tokens = exactly 3 query tiles, so here the three tiles of a (frame, head) item run side by side — one group each,
// 3 softmax warps per scheduler — on the same K / V, in steps of 64 keys:
//   TMEM  O_g [128 x 64] fp32 at 64 g;  S_g [128 x 64] fp32 at 192 + 64 g;  P_g [128 x 64] bf16 at 384 + 32 g   (480 columns)
//   smem  Q 3 groups x 2 slots, K ring of 4 chunks of 128 keys, V ring of 4 chunks   (14 x 16 KB = 224 KB)
//   warp 0 TMA producer; warps 1-3 MMA issuer of group 0-2 (warp 2 also allocates TMEM); warps 4-15 softmax + epilogue
// Per step and group: S = Q K_s^T (4 MMAs, N = 64) -> the row's 64 scores in registers (S is free again, the next S is
// issued) -> max, lazy raise of the reference (attn_spatial_pp.cuh), exp2, round to bf16 -> P (TMEM) -> O += P V_s.
// A K / V chunk is released when all tiles of the item are through both of its halves; with rings of 4 the next item's
// first two chunks are loaded while the current item is being worked on.
#pragma once

namespace istvt {

constexpr int S3_GROUPS = 3;
constexpr int S3_THREADS = 128 + 128 * S3_GROUPS;        // 512
constexpr int S3_RING = 4;                                // K / V chunks of 128 keys in flight
constexpr int S3_Q_OFF = 0;                               // [group][slot][16 KB]
constexpr int S3_K_OFF = 2 * S3_GROUPS * SP_CHUNK_BYTES;
constexpr int S3_V_OFF = S3_K_OFF + S3_RING * SP_CHUNK_BYTES;
constexpr int S3_MISC_OFF = S3_V_OFF + S3_RING * SP_CHUNK_BYTES;   // 224 KB
constexpr int S3_SMEM = S3_MISC_OFF + 1024 /*align*/ + 512 /*barriers*/;
constexpr int S3_TMEM_O = 0, S3_TMEM_S = 192, S3_TMEM_P = 384;

__global__ void __launch_bounds__(S3_THREADS, 1)
attn_spatial_pp3_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, int tokens,
                        int heads, int items, float scale_log2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_q = smem + S3_Q_OFF;
    uint8_t* s_k = smem + S3_K_OFF;
    uint8_t* s_v = smem + S3_V_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S3_MISC_OFF);
    uint64_t* q_full = bars;             // [3 groups x 2 slots]
    uint64_t* q_empty = bars + 6;        // [6]
    uint64_t* k_full = bars + 12;        // [4]
    uint64_t* k_empty = bars + 16;       // [4]   q_tiles arrivals: every tile of the item is through the chunk
    uint64_t* v_full = bars + 20;        // [4]
    uint64_t* v_empty = bars + 24;       // [4]   q_tiles arrivals
    uint64_t* s_full = bars + 28;        // [3]   S step of group g is in TMEM
    uint64_t* s_free = bars + 31;        // [3]   the group has it in registers (4 warps)
    uint64_t* p_full = bars + 34;        // [3]   P step written, O rescaled / drained as needed (4 warps)
    uint64_t* pv_done = bars + 37;       // [3]   PV of the group's previous step has retired
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 40);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int inner = heads * SA_DH;
    const int q_tiles = (tokens + SA_BM - 1) / SA_BM;
    const int k_chunks = q_tiles;
    const int steps = (tokens + 63) / 64;
    const int my_items = (items > static_cast<int>(blockIdx.x))
                             ? (items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                   static_cast<int>(gridDim.x)
                             : 0;
    const int n_tiles = my_items * q_tiles;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 6; ++i) { mbar_init(q_full + i, 1); mbar_init(q_empty + i, 1); }
        for (int i = 0; i < S3_RING; ++i) {
            mbar_init(k_full + i, 1); mbar_init(k_empty + i, q_tiles);
            mbar_init(v_full + i, 1); mbar_init(v_empty + i, q_tiles);
        }
        for (int i = 0; i < S3_GROUPS; ++i) {
            mbar_init(s_full + i, 1); mbar_init(s_free + i, 4);
            mbar_init(p_full + i, 4); mbar_init(pv_done + i, 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_holder, SA_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int n = 0; n < my_items; ++n) {
                const int item = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
                const int h = item % heads;
                const int bf = item / heads;
                for (int qt = 0; qt < q_tiles; ++qt) {
                    const int t = n * q_tiles + qt;
                    const int g = t % S3_GROUPS, cnt = t / S3_GROUPS;
                    const int qs = g * 2 + (cnt & 1);
                    mbar_wait_sleep(q_empty + qs, ((cnt >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(q_full + qs, SP_CHUNK_BYTES);
                    tma_load_3d(s_q + qs * SP_CHUNK_BYTES, &tm_qkv, q_full + qs, h * SA_DH, qt * SA_BM, bf);
                }
                for (int c = 0; c < k_chunks; ++c) {
                    const int i = n * k_chunks + c;
                    const int slot = i & (S3_RING - 1);
                    const uint32_t ph = ((i / S3_RING) & 1) ^ 1;
                    mbar_wait_sleep(k_empty + slot, ph);
                    mbar_arrive_expect_tx(k_full + slot, SP_CHUNK_BYTES);
                    tma_load_3d(s_k + slot * SP_CHUNK_BYTES, &tm_qkv, k_full + slot, inner + h * SA_DH, c * 128, bf);
                    mbar_wait_sleep(v_empty + slot, ph);
                    mbar_arrive_expect_tx(v_full + slot, SP_CHUNK_BYTES);
                    tma_load_3d(s_v + slot * SP_CHUNK_BYTES, &tm_qkv, v_full + slot, 2 * inner + h * SA_DH, c * 128, bf);
                }
            }
        }
        __syncwarp();
    } else if (warp < 4) {
        // ================= MMA issuer of group g =================
        const int g = warp - 1;
        if (elect_one()) {
            const uint32_t idesc_s = make_idesc_bf16(SA_BM, 64, 0, 0);
            const uint32_t idesc_pv = make_idesc_bf16(SA_BM, SA_DH, 0, 1);   // B (= V) is MN-major
            const uint64_t desc_kmaj = make_smem_desc(0, 0, 1024, SWZ_128B);
            const uint64_t desc_v = make_smem_desc(smem_u32(s_v), 64 * 128, 1024, SWZ_128B);
            const uint32_t q_field = (smem_u32(s_q) & 0x3FFFFu) >> 4;
            const uint32_t k_field = (smem_u32(s_k) & 0x3FFFFu) >> 4;
            const int last_ksteps = (tokens - (steps - 1) * 64 + 15) / 16;
            const uint32_t t_o = tmem_base + S3_TMEM_O + g * 64;
            const uint32_t t_s = tmem_base + S3_TMEM_S + g * 64;
            const uint32_t t_p = tmem_base + S3_TMEM_P + g * 32;
            int gidx = 0;                // steps issued by this group
            int prev_s = 0, prev_i = 0;  // the step whose PV is still to be issued

            auto issue_pv = [&](int s, int i, int pidx) {
                const int slot = i & (S3_RING - 1);
                if ((s & 1) == 0) mbar_wait_hot(v_full + slot, (i / S3_RING) & 1);
                mbar_wait_hot(p_full + g, pidx & 1);
                tc_fence_after();
                const uint64_t b0 = desc_v + static_cast<uint64_t>((slot * 8 + (s & 1) * 4) * (16 * 128 >> 4));
                const int ksteps = (s == steps - 1) ? last_ksteps : 4;
                for (int j = 0; j < ksteps; ++j)
                    umma_f16_ts(t_o, t_p + j * 8, b0 + j * (16 * 128 >> 4), idesc_pv, (s | j) != 0 ? 1u : 0u);
                umma_commit(pv_done + g);
                if ((s & 1) != 0 || s == steps - 1) umma_commit(v_empty + slot);
            };

            int n = 0, qt = g;           // tile t = n * q_tiles + qt
            while (qt >= q_tiles) { qt -= q_tiles; ++n; }
            int cnt = 0;
            for (int t = g; t < n_tiles; t += S3_GROUPS, ++cnt) {
                const int qs = g * 2 + (cnt & 1);
                const uint64_t q_desc = desc_kmaj | (q_field + qs * (SP_CHUNK_BYTES >> 4));
                for (int s = 0; s < steps; ++s) {
                    const int i = n * k_chunks + (s >> 1);
                    const int slot = i & (S3_RING - 1);
                    if (s == 0) mbar_wait_hot(q_full + qs, (cnt >> 1) & 1);
                    if ((s & 1) == 0) mbar_wait_hot(k_full + slot, (i / S3_RING) & 1);
                    if (gidx > 0) mbar_wait_hot(s_free + g, (gidx - 1) & 1);
                    tc_fence_after();
                    // 64 keys = rows [64 (s & 1), +64) of the 128-row K chunk: 8 swizzle atoms of 1 KB further on
                    const uint64_t kc = desc_kmaj | (k_field + slot * (SP_CHUNK_BYTES >> 4) + (s & 1) * (8192 >> 4));
                    umma_f16_ss(t_s, q_desc, kc, idesc_s, 0u);
                    umma_f16_ss(t_s, q_desc + 2, kc + 2, idesc_s, 1u);
                    umma_f16_ss(t_s, q_desc + 4, kc + 4, idesc_s, 1u);
                    umma_f16_ss(t_s, q_desc + 6, kc + 6, idesc_s, 1u);
                    umma_commit(s_full + g);
                    if (s == steps - 1) umma_commit(q_empty + qs);
                    if ((s & 1) != 0 || s == steps - 1) umma_commit(k_empty + slot);
                    if (gidx > 0) issue_pv(prev_s, prev_i, gidx - 1);
                    prev_s = s; prev_i = i;
                    ++gidx;
                }
                qt += S3_GROUPS;
                while (qt >= q_tiles) { qt -= q_tiles; ++n; }
            }
            if (gidx > 0) issue_pv(prev_s, prev_i, gidx - 1);
        }
        __syncwarp();
    } else {
        // ================= softmax + epilogue of group g, thread = query row =================
        const int g = (warp - 4) >> 2;
        const int quad = warp & 3;
        const int row = quad * 32 + lane;                 // row inside the q tile == TMEM lane
        const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t t_o = tmem_base + lane_base + S3_TMEM_O + g * 64;
        const uint32_t t_s = tmem_base + lane_base + S3_TMEM_S + g * 64;
        const uint32_t t_p = tmem_base + lane_base + S3_TMEM_P + g * 32;
        const f32x2_t sc2 = f32x2_make(scale_log2, scale_log2);
        float inv_prev = 0.0f;
        int out_prev = -1;                                 // element offset of this row's 64 outputs, -1 = no store
        int gidx = 0;

        auto epilogue = [&]() {                            // O of the previous tile -> global (its last PV has retired)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t ro[32];
                tmem_ld_32x32b_x32(t_o + half * 32, ro);
                tmem_ld_wait();
                if (out_prev >= 0) {
                    __nv_bfloat16* op = out + static_cast<size_t>(static_cast<uint32_t>(out_prev)) + half * 32;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(ro[8 * q + 0]) * inv_prev, __uint_as_float(ro[8 * q + 1]) * inv_prev);
                        o.y = pack_bf16x2(__uint_as_float(ro[8 * q + 2]) * inv_prev, __uint_as_float(ro[8 * q + 3]) * inv_prev);
                        o.z = pack_bf16x2(__uint_as_float(ro[8 * q + 4]) * inv_prev, __uint_as_float(ro[8 * q + 5]) * inv_prev);
                        o.w = pack_bf16x2(__uint_as_float(ro[8 * q + 6]) * inv_prev, __uint_as_float(ro[8 * q + 7]) * inv_prev);
                        *reinterpret_cast<uint4*>(op + 8 * q) = o;
                    }
                }
            }
        };

        int n = 0, qt = g;
        while (qt >= q_tiles) { qt -= q_tiles; ++n; }
        for (int t = g; t < n_tiles; t += S3_GROUPS) {
            float m_ref = 0.0f, l_run = 0.0f;
            for (int s = 0; s < steps; ++s, ++gidx) {
                const int vh = tokens - s * 64;            // valid columns of this step (warp-uniform, > 0)
                mbar_wait(s_full + g, gidx & 1);
                tc_fence_after();
                uint32_t r[2][32];
                tmem_ld_32x32b_x32(t_s, r[0]);
                if (vh > 32) tmem_ld_32x32b_x32(t_s + 32, r[1]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_free + g);    // the issuer may overwrite S with the next step

                // ---- maximum of the step; raise the reference only when it is exceeded by more than 2^8 ----
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int sb = 0; sb < 2; ++sb) {
                    const int vl = vh - sb * 32;
                    if (vl >= 32) s2_sub_max<false>(r[sb], vl, mx4);
                    else if (vl > 0) s2_sub_max<true>(r[sb], vl, mx4);
                }
                const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
                float corr = 1.0f;
                bool raise = false;
                if (s == 0) {
                    m_ref = mx;
                } else if ((mx - m_ref) * scale_log2 > S2_RESCALE_LOG2) {
                    corr = ex2_approx((m_ref - mx) * scale_log2);
                    m_ref = mx;
                    raise = true;
                }
                const bool any_raise = __any_sync(0xffffffffu, raise);
                const float mxs = m_ref * scale_log2;
                const f32x2_t nm2 = f32x2_make(-mxs, -mxs);

                // ---- P = exp2(S * scale - mxs) -> bf16 pairs -> TMEM, 32 keys at a time (the first store's latency hides
                //      behind the second half's exponentials); denominator from the unrounded fp32 values.  The group's
                //      previous PV (issued when the last step's P was complete) has had the load, the maximum and 32
                //      exponentials to retire before the first store needs the P columns. ----
                f32x2_t acc[2] = {0ull, 0ull};
#pragma unroll
                for (int sb = 0; sb < 2; ++sb) {
                    uint32_t pk[16];
                    const int vl = vh - sb * 32;
                    if (vl >= 32) {
                        s2_sub_exp<false, true>(r[sb], pk, vl, sc2, nm2, acc);
                    } else if (vl > 0) {
                        s2_sub_exp<true, true>(r[sb], pk, vl, sc2, nm2, acc);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[j] = 0u;
                    }
                    if (sb == 0 && gidx > 0) {
                        mbar_wait(pv_done + g, (gidx - 1) & 1);
                        tc_fence_after();
                    }
                    tmem_st_32x32b_x16(t_p + sb * 16, pk);
                }
                {
                    float a0, a1, a2, a3;
                    f32x2_split(acc[0], a0, a1);
                    f32x2_split(acc[1], a2, a3);
                    l_run = fmaf(l_run, corr, (a0 + a1) + (a2 + a3));
                }
                // ---- rare, warp-uniform: a row raised its reference -> O *= corr before this step's PV accumulates
                //      (here, where neither the scores nor P are live: two 32-register TMEM transfers) ----
                if (any_raise) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t ro[32];
                        tmem_ld_32x32b_x32(t_o + half * 32, ro);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * corr);
                        tmem_st_32x32b_x32(t_o + half * 32, ro);
                    }
                }
                if (s == 0 && t > g) epilogue();           // drain the previous tile's O before PV.0 overwrites it
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full + g);
            }
            {
                const int item = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
                const int h = item % heads;
                const int bf = item / heads;
                const int q_idx = qt * SA_BM + row;
                inv_prev = 1.0f / l_run;
                out_prev = (q_idx < tokens) ? (bf * tokens + q_idx) * inner + h * SA_DH : -1;
            }
            qt += S3_GROUPS;
            while (qt >= q_tiles) { qt -= q_tiles; ++n; }
        }
        if (gidx > 0) {
            mbar_wait(pv_done + g, (gidx - 1) & 1);
            tc_fence_after();
            epilogue();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, SA_TMEM_COLS);
    }
}

}  // namespace istvt
