// Joint attention, two query tiles per CTA in ping-pong (included by attn_joint.cu).
//
// Why: in attn_joint_tcgen05_kernel the softmax warps are the critical resource and run a latency chain per key block
// (tcgen05.ld -> max -> exchange -> ex2 -> st.shared -> fence -> mbarrier -> MMA) with both warps of a scheduler in the
// same phase: the XU pipe is busy 35 % of the time (profiles/r5d_attn_joint_ncu_source.txt).  Here one CTA owns TWO
// adjacent 128-query tiles of the same (sequence, head) — they share every K / V block, so each block is fetched once
// for 256 queries — and each tile has its own MMA-issuing thread, softmax group (4 warps, ONE thread per query row: no
// row-maximum exchange, no barrier inside a block), S / O accumulators in TMEM and P buffer.  The two groups drift into
// different phases, so a scheduler always holds one warp of each group: one can be in its MUFU phase while the other
// loads, stores or waits.
//
//   warp 0        TMA producer: Q_A, Q_B, then K / V blocks of 128 keys through a 3-stage ring (a stage is released
//                 when the PV MMAs of BOTH tiles have retired: kv_empty counts 2 commits)
//   warp 1 / 3    MMA issuer of tile A / B:  S_j = Q K_j^T -> TMEM (single buffer, re-issued as soon as the softmax
//                 group has read S_j for the second time), O_j = P_j V_j -> TMEM (accumulate = 0)
//   warps 4-7 / 8-11  softmax group of tile A / B, thread = query row:
//                 pass 1  max over the 128 scores (4 x tcgen05.ld of 32 columns)        -> m_new, corr
//                 fold    wait PV_{j-1}; o = o * corr_{j-1} + O_{j-1}  (64 register accumulators per row)
//                 pass 2  re-read S_j, p = exp2(s c - m c) -> bf16 -> P (smem, SW128), denominator l
//   TMEM : per tile S[128 x 128] + O[128 x 64] fp32 = 192 columns; 384 of 512 used.
//   smem : Q 2 x 16 KB, K/V ring 3 x 32 KB, P 2 x 32 KB = 192 KB.
#pragma once

namespace istvt {

constexpr int PP_STAGES = 3;
constexpr int PP_THREADS = 384;    // warps 0-3: TMA, MMA A, TMEM alloc, MMA B; warps 4-7: softmax A; warps 8-11: softmax B
constexpr int PP_SMEM = (2 + 2 * PP_STAGES) * JA_TILE_BYTES + 2 * JA_P_BYTES + 512 + 1024;
constexpr int PP_TMEM_COLS = 512;

template <bool ROUNDED_SUM>
__global__ void __launch_bounds__(PP_THREADS, 1)
attn_joint_pp_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, int tokens, int heads,
                     float scale_log2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_q = smem;                                   // [2 tiles][16 KB]
    uint8_t* s_k = s_q + 2 * JA_TILE_BYTES;                // [PP_STAGES][16 KB]
    uint8_t* s_v = s_k + PP_STAGES * JA_TILE_BYTES;        // [PP_STAGES][16 KB]
    uint8_t* s_p = s_v + PP_STAGES * JA_TILE_BYTES;        // [2 tiles][32 KB]
    uint8_t* s_misc = s_p + 2 * JA_P_BYTES;
    uint64_t* bar_q = reinterpret_cast<uint64_t*>(s_misc); // [2]
    uint64_t* kv_full = bar_q + 2;                         // [PP_STAGES]
    uint64_t* kv_empty = kv_full + PP_STAGES;              // [PP_STAGES], 2 arrivals (one per tile)
    uint64_t* bar_s = kv_empty + PP_STAGES;                // [2]  S_j of tile g is in TMEM
    uint64_t* bar_sfree = bar_s + 2;                       // [2]  the group has finished reading S_j
    uint64_t* bar_p = bar_sfree + 2;                       // [2]  P_j of tile g is in smem, O of tile g is drained
    uint64_t* bar_o = bar_p + 2;                           // [2]  PV_j of tile g has retired
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_o + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int q_tiles = (tokens + JA_BM - 1) / JA_BM;
    const int pairs = (q_tiles + 1) / 2;
    const int pr = blockIdx.x % pairs;
    const int h = (blockIdx.x / pairs) % heads;
    const int b = blockIdx.x / (pairs * heads);
    const int inner = heads * JA_DH;
    const int nblk = (tokens + JA_BN - 1) / JA_BN;
    const bool b_live = 2 * pr + 1 < q_tiles;              // an odd tile count leaves the last pair's second tile empty

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < PP_STAGES; ++s) {
            mbar_init(kv_full + s, 1);
            mbar_init(kv_empty + s, 2);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(bar_q + g, 1);
            mbar_init(bar_s + g, 1);
            mbar_init(bar_sfree + g, 4);   // one arrive per softmax warp of the group
            mbar_init(bar_p + g, 4);
            mbar_init(bar_o + g, 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_holder, PP_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        if (lane == 0) {
            for (int g = 0; g < 2; ++g) {
                mbar_arrive_expect_tx(bar_q + g, JA_TILE_BYTES);
                tma_load_3d(s_q + g * JA_TILE_BYTES, &tm_qkv, bar_q + g, h * JA_DH, (2 * pr + g) * JA_BM, b);
            }
            for (int j = 0; j < nblk; ++j) {
                const int s = j % PP_STAGES;
                if (j >= PP_STAGES) mbar_wait_sleep(kv_empty + s, ((j / PP_STAGES) - 1) & 1);
                mbar_arrive_expect_tx(kv_full + s, 2 * JA_TILE_BYTES);
                tma_load_3d(s_k + s * JA_TILE_BYTES, &tm_qkv, kv_full + s, inner + h * JA_DH, j * JA_BN, b);
                tma_load_3d(s_v + s * JA_TILE_BYTES, &tm_qkv, kv_full + s, 2 * inner + h * JA_DH, j * JA_BN, b);
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 3) {
        const int g = warp >> 1;                                         // tile A (warp 1) or B (warp 3)
        if (g == 1 && !b_live) {
            // empty tile: only keep the K / V ring's release count right
            for (int j = 0; j < nblk; ++j) {
                mbar_wait(kv_full + (j % PP_STAGES), (j / PP_STAGES) & 1);
                if (lane == 0) mbar_arrive(kv_empty + (j % PP_STAGES));
                __syncwarp();
            }
        } else {
            const uint32_t idesc_s = make_idesc_bf16(JA_BM, JA_BN, 0, 0);
            const uint32_t idesc_o = make_idesc_bf16(JA_BM, JA_DH, 0, 1);    // B (= V) is MN-major
            const uint32_t q_addr = smem_u32(s_q + g * JA_TILE_BYTES);
            const uint32_t p_addr = smem_u32(s_p + g * JA_P_BYTES);
            const uint32_t tmem_s = tmem_base + g * 192;
            const uint32_t tmem_o = tmem_s + JA_BN;
            mbar_wait(bar_q + g, 0);
            for (int j = 0; j <= nblk; ++j) {
                if (j < nblk) {   // S_j: needs K_j and (j > 0) the group to be done with S_{j-1}
                    const int st = j % PP_STAGES;
                    if (j > 0) mbar_wait(bar_sfree + g, (j - 1) & 1);
                    mbar_wait(kv_full + st, (j / PP_STAGES) & 1);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t k_addr = smem_u32(s_k + st * JA_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < JA_DH / 16; ++k)
                            umma_f16_ss(tmem_s, make_smem_desc(q_addr + k * 32, 0, 1024, SWZ_128B),
                                        make_smem_desc(k_addr + k * 32, 0, 1024, SWZ_128B), idesc_s, k != 0 ? 1u : 0u);
                        umma_commit(bar_s + g);
                    }
                    __syncwarp();
                }
                if (j > 0) {      // PV_{j-1}: P_{j-1} is in smem and the group has folded O_{j-2} away
                    const int jj = j - 1;
                    const int st = jj % PP_STAGES;
                    mbar_wait(bar_p + g, jj & 1);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t v_addr = smem_u32(s_v + st * JA_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < JA_BN / 16; ++k) {
                            const uint64_t a_desc =
                                make_smem_desc(p_addr + (k >> 2) * (JA_BM * 128) + (k & 3) * 32, 0, 1024, SWZ_128B);
                            const uint64_t b_desc = make_smem_desc(v_addr + k * 16 * 128, 64 * 128, 1024, SWZ_128B);
                            umma_f16_ss(tmem_o, a_desc, b_desc, idesc_o, k != 0 ? 1u : 0u);
                        }
                        umma_commit(bar_o + g);
                        umma_commit(kv_empty + st);   // this tile is done with K_jj / V_jj
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        const int g = (warp - 4) >> 2;
        const int quad = warp & 3;
        const int row = quad * 32 + lane;               // row inside the q tile == TMEM lane
        const int q_idx = (2 * pr + g) * JA_BM + row;
        if (g == 0 || b_live) {
            const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
            const uint32_t t_s = tmem_base + g * 192 + lane_base;
            const uint32_t t_o = t_s + JA_BN;
            uint8_t* prow0 = s_p + g * JA_P_BYTES + (row >> 3) * 1024 + (row & 7) * 128;

            float m_run = -INFINITY, l_run = 0.0f, corr_prev = 0.0f;
            float o[JA_DH];
#pragma unroll
            for (int i = 0; i < JA_DH; ++i) o[i] = 0.0f;

            auto fold = [&]() {                             // o = o * corr_prev + O
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t ro[32];
                    tmem_ld_32x32b_x32(t_o + half * 32, ro);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[half * 32 + i] = fmaf(o[half * 32 + i], corr_prev, __uint_as_float(ro[i]));
                }
            };

            for (int j = 0; j < nblk; ++j) {
                const int valid = tokens - j * JA_BN;       // columns [0, valid) of the block are real keys
                mbar_wait(bar_s + g, j & 1);
                tc_fence_after();
                // ---- pass 1: row maximum ----
                float mx = -INFINITY;
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    if (valid <= l * 32) break;             // warp-uniform
                    uint32_t r[1][32];
                    tmem_ld_32x32b_x32(t_s + l * 32, r[0]);
                    tmem_ld_wait();
                    const int vl = valid - l * 32;
                    mx = fmaxf(mx, vl < 32 ? block_max<1, true>(r, vl) : block_max<1, false>(r, vl));
                }
                const float m_new = fmaxf(m_run, mx);
                const float corr = ex2_approx((m_run - m_new) * scale_log2);
                const float mxs = m_new * scale_log2;
                // ---- PV_{j-1} has had pass 1 to retire: fold it in; that also frees the P buffer ----
                if (j > 0) {
                    mbar_wait(bar_o + g, (j - 1) & 1);
                    tc_fence_after();
                    fold();
                }
                // ---- pass 2: P_j ----
                float sum = 0.0f;
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    uint32_t r[1][32];
                    tmem_ld_32x32b_x32(t_s + l * 32, r[0]);
                    tmem_ld_wait();
                    if (l == 3) {                           // S_j is in registers: the issuer may overwrite it with S_{j+1}
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_sfree + g);
                    }
                    const int vl = valid - l * 32;
                    uint8_t* prow = prow0 + (l >> 1) * (JA_BM * 128);
                    const int chunk0 = (l & 1) * 4;
                    sum += vl < 32 ? block_exp_store<1, true, true, ROUNDED_SUM>(r, vl, scale_log2, mxs, prow, chunk0, row)
                                   : block_exp_store<1, false, true, ROUNDED_SUM>(r, vl, scale_log2, mxs, prow, chunk0, row);
                }
                l_run = fmaf(l_run, corr, sum);
                fence_proxy_async_smem();   // st.shared P -> visible to the tensor core (async proxy)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_p + g);
                corr_prev = corr;
                m_run = m_new;
            }
            mbar_wait(bar_o + g, (nblk - 1) & 1);
            tc_fence_after();
            fold();
            const float inv = 1.0f / l_run;
            if (q_idx < tokens) {
                __nv_bfloat16* op = out + (static_cast<int64_t>(b) * tokens + q_idx) * inner + h * JA_DH;
#pragma unroll
                for (int gq = 0; gq < JA_DH / 8; ++gq) {
                    uint4 v;
                    v.x = pack_bf16x2(o[8 * gq + 0] * inv, o[8 * gq + 1] * inv);
                    v.y = pack_bf16x2(o[8 * gq + 2] * inv, o[8 * gq + 3] * inv);
                    v.z = pack_bf16x2(o[8 * gq + 4] * inv, o[8 * gq + 5] * inv);
                    v.w = pack_bf16x2(o[8 * gq + 6] * inv, o[8 * gq + 7] * inv);
                    *reinterpret_cast<uint4*>(op + 8 * gq) = v;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, PP_TMEM_COLS);
    }
}

}  // namespace istvt
