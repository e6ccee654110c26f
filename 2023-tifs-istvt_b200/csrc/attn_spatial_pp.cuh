// Spatial attention, production kernel (included by attn_spatial.cu): persistent, two query tiles in flight per CTA,
// ONE thread per query row, online softmax with a lazy rescale.  network/vivit/module.py:84-91.
//
// Why (profiles/README.md r7k, r6u): in attn_spatial_pipe_kernel the 16 softmax warps walk every tile in lockstep
// (pass 1 max -> bar.sync -> pass 2 exp -> bar.sync), so the MUFU pipe — the bound of this op, one ex2 per score at
// 16 / clk / SM — idles during pass 1, the barriers, the deferred epilogue and the PV -> S hand-off (S[128 x 384] fills
// TMEM, the next tile's S cannot start before P has been consumed): 7.3 kclk per tile against 2.9 kclk of MUFU work, at
// ~10 issued instructions per score.  (tcgen05.ld is NOT the limit: 880-920 B/clk/SM measured, tools/micro/tmem_bw.cu.)
// Here
//   * a tile is processed in 128-key chunks: S chunk [128 x 128] fp32, P chunk [128 x 128] bf16 and O [128 x 64] fp32
//     take 256 TMEM columns, so TWO tiles (groups A / B, alternate tiles of the CTA's item walk) live side by side, each
//     with its own MMA-issuing thread and 4 softmax warps; the groups drift apart, one is in its ex2 phase while the
//     other loads, stores or waits;
//   * thread = query row: no row-maximum exchange, no block barrier.  The running maximum is only raised when a chunk
//     exceeds it by more than 2^8 (then O and the denominator are rescaled — with 3 chunks per row that is rare);
//   * S chunk c+1 is issued as soon as the group has chunk c in registers (P has its own columns), so S is always
//     waiting for the softmax warps, not the other way round;
//   * 3-4 issued instructions per score: FMNMX3 (two scores), FFMA2, 2 x MUFU.EX2, FADD2 and the bf16 pack per PAIR.
//     The pack rounds to nearest with integer adds + PRMT (F2FP would sit on the XU pipe with the exponentials).  A
//     variant that rounds for free (ISTVT_SA_ROUND=trunc: exponent biased by log2(1 + 0.002707), PRMT truncates, the
//     factor taken out of the final 1 / l; error within (-0.65, +0.69) ulp, mean 0) measured the same speed — issue
//     slots are not the limit — so round-to-nearest is the default.
//
//   warp 0        TMA producer: K (double buffered per item), Q (2 slots per group), V (one buffer, released per chunk)
//   warp 1 / 3    MMA issuer of group A / B:  S.c = Q K_c^T,  O (+)= P.c V_c (A operand from TMEM)
//   warp 2        TMEM allocator
//   warps 4-7 / 8-11  softmax + epilogue of group A / B
//   TMEM (per group, 256 columns): S [0,128)  P [128,192)  O [192,256)
//   smem: Q 4 x 16 KB, K 2 x 48 KB, V 48 KB = 208 KB
#pragma once

namespace istvt {

constexpr int S2_THREADS = 384;
constexpr int S2_Q_OFF = 0;
constexpr int S2_K_OFF = 4 * SP_CHUNK_BYTES;
constexpr int S2_V_OFF = S2_K_OFF + 2 * SA_KV_BYTES;
constexpr int S2_MISC_OFF = S2_V_OFF + SA_KV_BYTES;
constexpr int S2_SMEM = S2_MISC_OFF + 1024 /*align*/ + 512 /*barriers*/;
constexpr float S2_RESCALE_LOG2 = 8.0f;                 // raise the running maximum only beyond 2^8
constexpr float S2_ROUND_BIAS = 1.002707f;
constexpr float S2_ROUND_BIAS_LOG2 = 0.0039001f;        // log2(1 + 0.002707): truncation to bf16 becomes mean-zero rounding

// scores of one 32-column sub-block -> four independent running maxima (a single chain is 16 dependent FMNMX3)
template <bool MASK>
__device__ __forceinline__ void s2_sub_max(const uint32_t* r, int vl, float (&mx)[4]) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        float a = __uint_as_float(r[j]), b = __uint_as_float(r[j + 1]);
        if (MASK) {
            if (j >= vl) a = -INFINITY;
            if (j + 1 >= vl) b = -INFINITY;
        }
        mx[(j >> 1) & 3] = fmaxf(mx[(j >> 1) & 3], fmaxf(a, b));
    }
}

// 32 scores -> 16 packed bf16 pairs of exp2(s * scale - mxs); the unrounded values are added to acc (two pair sums)
template <bool MASK, bool RNE>
__device__ __forceinline__ void s2_sub_exp(const uint32_t* r, uint32_t* pk, int vl, f32x2_t sc2, f32x2_t nm2,
                                           f32x2_t (&acc)[2]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const f32x2_t x = f32x2_fma(f32x2_make_bits(r[2 * j], r[2 * j + 1]), sc2, nm2);
        float x0, x1;
        f32x2_split(x, x0, x1);
        float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
        if (MASK) {
            if (2 * j >= vl) e0 = 0.0f;
            if (2 * j + 1 >= vl) e1 = 0.0f;
        }
        acc[j & 1] = f32x2_add(acc[j & 1], f32x2_make(e0, e1));
        pk[j] = RNE ? pack_bf16x2_rne_alu(e0, e1) : __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632);
    }
}

template <bool RNE>
__global__ void __launch_bounds__(S2_THREADS, 1)
attn_spatial_pp_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
                       float* __restrict__ lse, int tokens, int heads, int items, float scale_log2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_q = smem + S2_Q_OFF;             // [group][slot][16 KB]
    uint8_t* s_k = smem + S2_K_OFF;             // [item parity][chunk][16 KB]
    uint8_t* s_v = smem + S2_V_OFF;             // [chunk][16 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S2_MISC_OFF);
    uint64_t* q_full = bars;             // [2 groups x 2 slots]
    uint64_t* q_empty = bars + 4;        // [4]
    uint64_t* k_full = bars + 8;         // [2]
    uint64_t* k_empty = bars + 10;       // [2]   q_tiles arrivals: the last S chunk of every tile of the item
    uint64_t* v_full = bars + 12;        // [3]   per chunk
    uint64_t* v_empty = bars + 15;       // [3]   q_tiles arrivals: PV.c of every tile of the item
    uint64_t* s_full = bars + 18;        // [2]   S chunk of group g is in TMEM
    uint64_t* s_free = bars + 20;        // [2]   the group has it in registers (4 warps)
    uint64_t* p_full = bars + 22;        // [2]   P chunk written, O rescaled / drained as needed (4 warps)
    uint64_t* pv_done = bars + 24;       // [2]   PV of the group's previous chunk has retired
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 26);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int inner = heads * SA_DH;
    const int q_tiles = (tokens + SA_BM - 1) / SA_BM;
    const int k_chunks = q_tiles;
    const int my_items = (items > static_cast<int>(blockIdx.x))
                             ? (items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                   static_cast<int>(gridDim.x)
                             : 0;
    const int n_tiles = my_items * q_tiles;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(q_full + i, 1); mbar_init(q_empty + i, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(k_full + i, 1); mbar_init(k_empty + i, q_tiles);
            mbar_init(s_full + i, 1); mbar_init(s_free + i, 4);
            mbar_init(p_full + i, 4); mbar_init(pv_done + i, 1);
        }
        for (int i = 0; i < 3; ++i) { mbar_init(v_full + i, 1); mbar_init(v_empty + i, q_tiles); }
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
                const int kb = n & 1;
                mbar_wait_sleep(k_empty + kb, ((n >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(k_full + kb, k_chunks * SP_CHUNK_BYTES);
                for (int c = 0; c < k_chunks; ++c)
                    tma_load_3d(s_k + kb * SA_KV_BYTES + c * SP_CHUNK_BYTES, &tm_qkv, k_full + kb, inner + h * SA_DH,
                                c * 128, bf);
                auto load_q = [&](int qt) {
                    const int t = n * q_tiles + qt;
                    const int qs = (t & 1) * 2 + ((t >> 1) & 1);        // group, slot of the group's tile counter
                    mbar_wait_sleep(q_empty + qs, ((t >> 2) & 1) ^ 1);
                    mbar_arrive_expect_tx(q_full + qs, SP_CHUNK_BYTES);
                    tma_load_3d(s_q + qs * SP_CHUNK_BYTES, &tm_qkv, q_full + qs, h * SA_DH, qt * SA_BM, bf);
                };
                // the first two tiles' Q before V (V waits for the previous item's PVs), the rest after
                for (int qt = 0; qt < q_tiles && qt < 2; ++qt) load_q(qt);
                for (int c = 0; c < k_chunks; ++c) {
                    mbar_wait_sleep(v_empty + c, (n & 1) ^ 1);
                    mbar_arrive_expect_tx(v_full + c, SP_CHUNK_BYTES);
                    tma_load_3d(s_v + c * SP_CHUNK_BYTES, &tm_qkv, v_full + c, 2 * inner + h * SA_DH, c * 128, bf);
                }
                for (int qt = 2; qt < q_tiles; ++qt) load_q(qt);
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 3) {
        // ================= MMA issuer of group g =================
        const int g = warp >> 1;
        if (elect_one()) {
            const uint32_t idesc_s = make_idesc_bf16(SA_BM, 128, 0, 0);
            const uint32_t idesc_pv = make_idesc_bf16(SA_BM, SA_DH, 0, 1);   // B (= V) is MN-major
            const uint64_t desc_kmaj = make_smem_desc(0, 0, 1024, SWZ_128B);
            const uint64_t desc_v = make_smem_desc(smem_u32(s_v), 64 * 128, 1024, SWZ_128B);
            const uint32_t q_field = (smem_u32(s_q) & 0x3FFFFu) >> 4;
            const uint32_t k_field = (smem_u32(s_k) & 0x3FFFFu) >> 4;
            const int last_ksteps = (tokens - (k_chunks - 1) * 128 + 15) / 16;
            const uint32_t t_s = tmem_base + g * 256;
            const uint32_t t_p = t_s + 128;
            const uint32_t t_o = t_s + 192;
            int sidx = 0;                // chunks issued by this group
            int prev_c = 0, prev_n = 0;  // the chunk whose PV is still to be issued

            auto issue_pv = [&](int c, int n, int pidx) {
                mbar_wait_hot(v_full + c, n & 1);
                mbar_wait_hot(p_full + g, pidx & 1);
                tc_fence_after();
                const uint64_t b0 = desc_v + static_cast<uint64_t>(c * 8 * (16 * 128 >> 4));
                const int ksteps = (c == k_chunks - 1) ? last_ksteps : 8;
                for (int j = 0; j < ksteps; ++j)
                    umma_f16_ts(t_o, t_p + j * 8, b0 + j * (16 * 128 >> 4), idesc_pv, (c | j) != 0 ? 1u : 0u);
                umma_commit(pv_done + g);
                umma_commit(v_empty + c);
            };

            int n = 0, qt = g;           // tile t = n * q_tiles + qt
            while (qt >= q_tiles) { qt -= q_tiles; ++n; }
            for (int t = g; t < n_tiles; t += 2) {
                const int cnt = t >> 1;
                const int qs = g * 2 + (cnt & 1);
                const int kb = n & 1;
                const uint64_t q_desc = desc_kmaj | (q_field + qs * (SP_CHUNK_BYTES >> 4));
                const uint64_t k_desc = desc_kmaj | (k_field + kb * (SA_KV_BYTES >> 4));
                for (int c = 0; c < k_chunks; ++c) {
                    if (c == 0) {
                        mbar_wait_hot(q_full + qs, (cnt >> 1) & 1);
                        mbar_wait_hot(k_full + kb, (n >> 1) & 1);
                    }
                    if (sidx > 0) mbar_wait_hot(s_free + g, (sidx - 1) & 1);
                    tc_fence_after();
                    const uint64_t kc = k_desc + static_cast<uint64_t>(c * (SP_CHUNK_BYTES >> 4));
                    umma_f16_ss(t_s, q_desc, kc, idesc_s, 0u);
                    umma_f16_ss(t_s, q_desc + 2, kc + 2, idesc_s, 1u);
                    umma_f16_ss(t_s, q_desc + 4, kc + 4, idesc_s, 1u);
                    umma_f16_ss(t_s, q_desc + 6, kc + 6, idesc_s, 1u);
                    umma_commit(s_full + g);
                    if (c == k_chunks - 1) {
                        umma_commit(q_empty + qs);
                        umma_commit(k_empty + kb);
                    }
                    if (sidx > 0) issue_pv(prev_c, prev_n, sidx - 1);
                    prev_c = c; prev_n = n;
                    ++sidx;
                }
                qt += 2;
                while (qt >= q_tiles) { qt -= q_tiles; ++n; }
            }
            if (sidx > 0) issue_pv(prev_c, prev_n, sidx - 1);
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ================= softmax + epilogue of group g, thread = query row =================
        const int g = (warp - 4) >> 2;
        const int quad = warp & 3;
        const int row = quad * 32 + lane;                 // row inside the q tile == TMEM lane
        const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t t_s = tmem_base + g * 256 + lane_base;
        const uint32_t t_p = t_s + 128;
        const uint32_t t_o = t_s + 192;
        const f32x2_t sc2 = f32x2_make(scale_log2, scale_log2);
        float inv_prev = 0.0f;
        int64_t out_prev = -1;                             // element offset of this row's 64 outputs, -1 = no store
        int sidx = 0;

        auto epilogue = [&]() {                            // O of the previous tile -> global (its last PV has retired)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t ro[32];
                tmem_ld_32x32b_x32(t_o + half * 32, ro);
                tmem_ld_wait();
                if (out_prev >= 0) {
                    __nv_bfloat16* op = out + out_prev + half * 32;
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
        for (int t = g; t < n_tiles; t += 2) {
            const int item = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
            const int h = item % heads;
            const int bf = item / heads;
            const int q_idx = qt * SA_BM + row;
            float m_ref = 0.0f, l_run = 0.0f;

            for (int c = 0; c < k_chunks; ++c, ++sidx) {
                const int valid = tokens - c * 128;        // columns [0, valid) of the chunk are real keys
                mbar_wait(s_full + g, sidx & 1);
                tc_fence_after();
                // The chunk goes through the registers in two halves of 64 scores (a whole chunk per thread spills);
                // each half is one step of the online softmax.
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    const int vh = valid - hb * 64;        // valid columns of this half (warp-uniform)
                    uint32_t r[2][32];
                    if (vh > 0) tmem_ld_32x32b_x32(t_s + hb * 64, r[0]);
                    if (vh > 32) tmem_ld_32x32b_x32(t_s + hb * 64 + 32, r[1]);
                    tmem_ld_wait();
                    if (hb == 1) {                         // S is in registers: the issuer may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(s_free + g);
                    }
                    // ---- maximum of the half; raise the reference only when it is exceeded by more than 2^8 ----
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
                    if (c == 0 && hb == 0) {
                        m_ref = mx;
                    } else if ((mx - m_ref) * scale_log2 > S2_RESCALE_LOG2) {
                        corr = ex2_approx((m_ref - mx) * scale_log2);
                        m_ref = mx;
                        raise = true;
                    }
                    const bool any_raise = __any_sync(0xffffffffu, raise);
                    const float mxs = fmaf(m_ref, scale_log2, -(RNE ? 0.0f : S2_ROUND_BIAS_LOG2));
                    const f32x2_t nm2 = f32x2_make(-mxs, -mxs);

                    // ---- rare, warp-uniform: a row raised its reference -> rescale what was accumulated so far.  Needs the
                    //      group's previous PV retired (O consistent); done before the exponentials so that the 32-register
                    //      TMEM transfers do not collide with the packed P values ----
                    bool pv_waited = false;
                    if (any_raise) {
                        if (hb == 0 && sidx > 0) {
                            mbar_wait(pv_done + g, (sidx - 1) & 1);
                            tc_fence_after();
                            pv_waited = true;
                        }
                        if (c > 0) {                       // O of this tile (for c == 0 it still holds the previous tile)
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
                        if (hb == 1) {                     // the first half of this chunk's P
                            tmem_st_wait();
                            uint32_t rp[32];
                            tmem_ld_32x32b_x32(t_p, rp);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const float2 v = unpack_bf16x2(rp[i]);
                                rp[i] = pack_bf16x2(v.x * corr, v.y * corr);
                            }
                            tmem_st_32x32b_x32(t_p, rp);
                        }
                    }

                    // ---- P = exp2(S * scale - mxs) -> bf16 pairs; denominator from the same fp32 values ----
                    f32x2_t acc[2] = {0ull, 0ull};
                    uint32_t pk[2][16];
#pragma unroll
                    for (int sb = 0; sb < 2; ++sb) {
                        const int vl = vh - sb * 32;
                        if (vl >= 32) {
                            s2_sub_exp<false, RNE>(r[sb], pk[sb], vl, sc2, nm2, acc);
                        } else if (vl > 0) {
                            s2_sub_exp<true, RNE>(r[sb], pk[sb], vl, sc2, nm2, acc);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) pk[sb][j] = 0u;
                        }
                    }
                    {
                        float a0, a1, a2, a3;
                        f32x2_split(acc[0], a0, a1);
                        f32x2_split(acc[1], a2, a3);
                        l_run = fmaf(l_run, corr, (a0 + a1) + (a2 + a3));
                    }
                    // ---- the previous PV of this group (issued when the last chunk's P was complete) has had the first
                    //      half's exponentials to retire: the P columns are free ----
                    if (hb == 0 && sidx > 0 && !pv_waited) {
                        mbar_wait(pv_done + g, (sidx - 1) & 1);
                        tc_fence_after();
                    }
                    tmem_st_32x32b_x16(t_p + hb * 32, pk[0]);
                    tmem_st_32x32b_x16(t_p + hb * 32 + 16, pk[1]);
                }
                if (c == 0 && t > g) epilogue();           // drain the previous tile's O before PV.0 overwrites it
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full + g);
            }
            // truncated P estimates the UNBIASED exponentials while l_run sums the biased ones: take the factor out again
            inv_prev = (RNE ? 1.0f : S2_ROUND_BIAS) / l_run;
            out_prev = (q_idx < tokens) ? (static_cast<int64_t>(bf) * tokens + q_idx) * inner + h * SA_DH : -1;
            if (lse != nullptr && q_idx < tokens)
                lse[(static_cast<int64_t>(bf) * heads + h) * tokens + q_idx] =
                    fmaf(m_ref, scale_log2, -(RNE ? 0.0f : S2_ROUND_BIAS_LOG2)) + log2f(l_run);
            qt += 2;
            while (qt >= q_tiles) { qt -= q_tiles; ++n; }
        }
        if (sidx > 0) {
            mbar_wait(pv_done + g, (sidx - 1) & 1);
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
