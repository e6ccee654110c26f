// Spatial self-attention over the tokens of one frame (network/vivit/module.py:84-91).
//
// bf16 path (tcgen05): one CTA = one (frame bf, head h, 128-query tile).  Q, K, V tiles are fetched by
// TMA straight out of the packed projection output [rows, 3*heads*64] through a 3-D tensor map
// (col, token-in-frame, frame), so rows past the frame's last token are zero-filled instead of
// leaking the next frame — no permute copies (the reference makes 3 in, 1 out).
//   S[128 x 384] = Q Kᵀ          tcgen05.mma, both operands K-major SW128, accumulator in TMEM (384 cols)
//   P = exp2(S*c - max*c)        8 softmax warps, 2 threads per row (192 columns each), whole key
//                                range resident in TMEM so no online rescaling is needed
//   O[128 x 64] = P V            P staged in smem (bf16, K-major SW128), V is the MN-major B operand
//   out = O / rowsum             TMEM -> registers -> global (128 B per row)
// fp32 path: SIMT validation kernel (online softmax, K/V in smem).
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "ptx.cuh"
#include "simt_util.cuh"

namespace istvt {

constexpr int SA_DH = 64;
constexpr int SA_BM = 128;        // queries per CTA
constexpr int SA_KMAX = 384;      // keys held in TMEM / smem
constexpr int SA_THREADS = 384;   // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: softmax + epilogue
constexpr int SA_Q_BYTES = SA_BM * SA_DH * 2;        // 16 KB
constexpr int SA_KV_BYTES = SA_KMAX * SA_DH * 2;     // 48 KB
constexpr int SA_P_BYTES = SA_BM * SA_KMAX * 2;      // 96 KB
constexpr int SA_SMEM = SA_Q_BYTES + 2 * SA_KV_BYTES + SA_P_BYTES + 1024 + 2048;
constexpr int SA_TMEM_COLS = 512;                     // S: 384, O: 64 (at column 384)

__global__ void __launch_bounds__(SA_THREADS, 1)
attn_spatial_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
                            float* __restrict__ probs, int tokens, int heads, float scale_log2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_q = smem;
    uint8_t* s_k = s_q + SA_Q_BYTES;
    uint8_t* s_v = s_k + SA_KV_BYTES;
    uint8_t* s_p = s_v + SA_KV_BYTES;
    uint8_t* s_misc = s_p + SA_P_BYTES;
    uint64_t* bar_qk = reinterpret_cast<uint64_t*>(s_misc);
    uint64_t* bar_v = bar_qk + 1;
    uint64_t* bar_s = bar_qk + 2;
    uint64_t* bar_p = bar_qk + 3;
    uint64_t* bar_o = bar_qk + 4;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_qk + 5);
    float* s_red = reinterpret_cast<float*>(s_misc + 64);  // [2][128] partial max, then partial sums

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int q_tiles = (tokens + SA_BM - 1) / SA_BM;
    const int qt = blockIdx.x % q_tiles;
    const int h = (blockIdx.x / q_tiles) % heads;
    const int bf = blockIdx.x / (q_tiles * heads);
    const int inner = heads * SA_DH;
    const int k_chunks = (tokens + 127) / 128;        // 128-key TMA boxes / MMA N-chunks
    const int keys_pad = k_chunks * 128;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
    if (warp == 1 && lane == 0) {
        mbar_init(bar_qk, 1);
        mbar_init(bar_v, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, 8);   // one arrive per softmax warp
        mbar_init(bar_o, 1);
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
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + SA_KMAX;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_qk, SA_Q_BYTES + k_chunks * 128 * SA_DH * 2);
            tma_load_3d(s_q, &tm_qkv, bar_qk, h * SA_DH, qt * SA_BM, bf);
            for (int c = 0; c < k_chunks; ++c)
                tma_load_3d(s_k + c * 128 * SA_DH * 2, &tm_qkv, bar_qk, inner + h * SA_DH, c * 128, bf);
            mbar_arrive_expect_tx(bar_v, k_chunks * 128 * SA_DH * 2);
            for (int c = 0; c < k_chunks; ++c)
                tma_load_3d(s_v + c * 128 * SA_DH * 2, &tm_qkv, bar_v, 2 * inner + h * SA_DH, c * 128, bf);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---- S = Q K^T ----
        mbar_wait(bar_qk, 0);
        tc_fence_after();
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(SA_BM, 128, 0, 0);
            const uint32_t q_addr = smem_u32(s_q);
            const uint32_t k_addr = smem_u32(s_k);
            for (int c = 0; c < k_chunks; ++c) {
#pragma unroll
                for (int k = 0; k < SA_DH / 16; ++k) {
                    const uint64_t a_desc = make_smem_desc(q_addr + k * 32, 0, 1024, SWZ_128B);
                    const uint64_t b_desc = make_smem_desc(k_addr + c * 128 * 128 + k * 32, 0, 1024, SWZ_128B);
                    umma_f16_ss(tmem_s + c * 128, a_desc, b_desc, idesc, k != 0 ? 1u : 0u);
                }
            }
            umma_commit(bar_s);
        }
        __syncwarp();
        // ---- O = P V ----
        mbar_wait(bar_v, 0);
        mbar_wait(bar_p, 0);
        tc_fence_after();
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(SA_BM, SA_DH, 0, 1);  // B (= V) is MN-major
            const uint32_t p_addr = smem_u32(s_p);
            const uint32_t v_addr = smem_u32(s_v);
            const int ksteps = (tokens + 15) / 16;
            for (int k = 0; k < ksteps; ++k) {
                const uint64_t a_desc =
                    make_smem_desc(p_addr + (k >> 2) * (SA_BM * 128) + (k & 3) * 32, 0, 1024, SWZ_128B);
                const uint64_t b_desc = make_smem_desc(v_addr + k * 16 * 128, 64 * 128, 1024, SWZ_128B);
                umma_f16_ss(tmem_o, a_desc, b_desc, idesc, k != 0 ? 1u : 0u);
            }
            umma_commit(bar_o);
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ---- softmax + epilogue: thread -> (row, column half) ----
        const int quad = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = quad * 32 + lane;               // row inside the q tile == TMEM lane
        const int q_idx = qt * SA_BM + row;             // token index of this query
        const int cols_half = keys_pad / 2;             // 192 (or 128 / 64 for short frames)
        const int col_lo = half * cols_half;
        const uint32_t t_row = tmem_s + (static_cast<uint32_t>(quad * 32) << 16);

        mbar_wait(bar_s, 0);
        tc_fence_after();

        // pass 1: row max over valid keys
        float mx = -INFINITY;
        for (int c0 = col_lo; c0 < col_lo + cols_half; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_row + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < tokens) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        s_red[half * 128 + row] = mx;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mx = fmaxf(s_red[row], s_red[128 + row]);
        const float mxs = mx * scale_log2;
        asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone has read the maxima; s_red reused for sums

        // pass 2: p = exp2(s*c - max*c), bf16 P -> smem (K-major SW128 atoms of 64 keys), row sums
        float sum = 0.0f;
        for (int c0 = col_lo; c0 < col_lo + cols_half; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_row + c0, r);
            tmem_ld_wait();
            float pv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float e = exp2f(fmaf(__uint_as_float(r[j]), scale_log2, -mxs));
                pv[j] = (c0 + j < tokens) ? e : 0.0f;
            }
            // The PV MMA consumes bf16 P; accumulate the denominator from the same rounded values so
            // that the normalised rows sum to one in the precision actually used.
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                pk[j] = pack_bf16x2(pv[2 * j], pv[2 * j + 1]);
                const float2 f = unpack_bf16x2(pk[j]);
                sum += f.x + f.y;
            }
            const int atom = c0 >> 6;                   // 64-key atom
            const int chunk0 = (c0 & 63) >> 3;          // first 16-byte chunk inside the 128-byte row
            uint8_t* prow = s_p + atom * (SA_BM * 128) + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int chunk = (chunk0 + g) ^ (row & 7);
                *reinterpret_cast<uint4*>(prow + chunk * 16) =
                    make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
            }
        }
        fence_proxy_async_smem();  // st.shared P -> visible to the tensor core (async proxy)
        s_red[half * 128 + row] = sum;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float inv = 1.0f / (s_red[row] + s_red[128 + row]);

        if (probs != nullptr) {  // warp-uniform
            // optional attention-map output: probs[bf, h, q, key], fp32, normalised.  The TMEM loads are
            // warp-collective, so rows past the frame end take part in them and only skip the stores.
            const bool row_ok = q_idx < tokens;
            float* pr = probs + ((static_cast<int64_t>(bf) * heads + h) * tokens + (row_ok ? q_idx : 0)) * tokens;
            for (int c0 = col_lo; c0 < col_lo + cols_half; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_row + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (row_ok && c0 + j < tokens)
                        pr[c0 + j] = exp2f(fmaf(__uint_as_float(r[j]), scale_log2, -mxs)) * inv;
            }
        }

        // epilogue: this thread normalises 32 of the 64 output columns of its row
        mbar_wait(bar_o, 0);
        tc_fence_after();
        {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_o + (static_cast<uint32_t>(quad * 32) << 16) + half * 32, r);
            tmem_ld_wait();
            if (q_idx < tokens) {
                __nv_bfloat16* op = out + (static_cast<int64_t>(bf) * tokens + q_idx) * inner + h * SA_DH + half * 32;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(r[8 * g + 0]) * inv, __uint_as_float(r[8 * g + 1]) * inv);
                    o.y = pack_bf16x2(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv);
                    o.z = pack_bf16x2(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv);
                    o.w = pack_bf16x2(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv);
                    *reinterpret_cast<uint4*>(op + 8 * g) = o;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, SA_TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// bf16 production kernel: persistent, software-pipelined.
//
// One CTA per SM walks (frame, head) items; an item is q_tiles (3) query tiles of 128 against the
// same K / V (3 chunks of 128 keys).  Warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4-19 = softmax + epilogue (4 threads per query row, 32 columns of every chunk each:
// 4 softmax warps per scheduler — with 2 the issue slots were 42 % used, profiles/r1o).
//   TMEM : S[128 x 384] fp32 (cols 0-383), O0 / O1 [128 x 64] fp32 (cols 384-447 / 448-511).
//   smem : Q ring 4 x 16 KB, K double buffer 2 x 48 KB (the next item's K and Q are prefetched),
//          V 48 KB (reloaded as soon as the item's last PV retires).
// P never touches shared memory: the softmax warps overwrite the S columns they have just consumed
// with bf16 P (tcgen05.st, two keys per 32-bit column) and the PV MMA takes its A operand from TMEM.
// Chunk-level hand-off keeps the tensor pipe and the MUFU pipe busy at the same time: the issuer
// interleaves PV(t-1).c with S(t).c as soon as the softmax warps release chunk c, so S(t) is ready
// when they finish tile t-1, and the O epilogue of tile t-1 is deferred until after the max pass of
// tile t (O is double buffered).  Bound: MUFU (ex2) — 128 x 362 exponentials per tile at 16/clk/SM.
// Optional `lse` output (training): log2-domain log-sum-exp of every query row, consumed by the backward.
// ------------------------------------------------------------------------------------------
constexpr int SP_QSLOTS = 4;
constexpr int SP_PARTS = 4;                                      // softmax threads per query row
constexpr int SP_SM_WARPS = 4 * SP_PARTS;                        // 16 softmax warps
constexpr int SP_THREADS = 128 + 32 * SP_SM_WARPS;               // 640
constexpr int SP_CHUNK_BYTES = 128 * SA_DH * 2;                  // 16 KB: 128 rows x 64 bf16
constexpr int SP_Q_OFF = 0;
constexpr int SP_K_OFF = SP_QSLOTS * SP_CHUNK_BYTES;              // 64 KB
constexpr int SP_V_OFF = SP_K_OFF + 2 * SA_KV_BYTES;              // + 96 KB
constexpr int SP_MISC_OFF = SP_V_OFF + SA_KV_BYTES;               // + 48 KB = 208 KB
constexpr int SP_SMEM = SP_MISC_OFF + 1024 /*align*/ + 256 /*barriers*/ + 2 * SP_PARTS * 128 * 4 /*row max, row sum*/;

__global__ void __launch_bounds__(SP_THREADS, 1)
attn_spatial_pipe_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
                         float* __restrict__ lse, int tokens, int heads, int items, float scale_log2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_q = smem + SP_Q_OFF;
    uint8_t* s_k = smem + SP_K_OFF;
    uint8_t* s_v = smem + SP_V_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SP_MISC_OFF);
    uint64_t* q_full = bars;            // [4]
    uint64_t* q_empty = bars + 4;       // [4]
    uint64_t* k_full = bars + 8;        // [2]
    uint64_t* k_empty = bars + 10;      // [2]
    uint64_t* v_full = bars + 12;
    uint64_t* v_empty = bars + 13;
    uint64_t* s_full = bars + 14;       // [3] S chunk c of the current tile is in TMEM
    uint64_t* p_full = bars + 17;       // [3] P chunk c written (and S chunk c consumed) by all softmax warps
    uint64_t* o_full = bars + 20;       // [2]
    uint64_t* o_empty = bars + 22;      // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 24);
    float* s_max = reinterpret_cast<float*>(smem + SP_MISC_OFF + 256);   // [SP_PARTS][128]
    float* s_sum = s_max + SP_PARTS * 128;                                // [SP_PARTS][128]

    const int warp = threadIdx.x >> 5;   // (roles in the four highest warp ids instead: neutral, 0.286 ms both ways, r6o)
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
            mbar_init(k_full + i, 1); mbar_init(k_empty + i, 1);
            mbar_init(o_full + i, 1); mbar_init(o_empty + i, SP_SM_WARPS);
        }
        mbar_init(v_full, 1); mbar_init(v_empty, 1);
        for (int i = 0; i < 3; ++i) { mbar_init(s_full + i, 1); mbar_init(p_full + i, SP_SM_WARPS); }
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
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + SA_KMAX;    // + 64 * buffer

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
                for (int qt = 0; qt < q_tiles; ++qt) {
                    const int t = n * q_tiles + qt;
                    const int slot = t & 3;
                    mbar_wait_sleep(q_empty + slot, ((t >> 2) & 1) ^ 1);
                    mbar_arrive_expect_tx(q_full + slot, SP_CHUNK_BYTES);
                    tma_load_3d(s_q + slot * SP_CHUNK_BYTES, &tm_qkv, q_full + slot, h * SA_DH, qt * SA_BM, bf);
                }
                mbar_wait_sleep(v_empty, (n & 1) ^ 1);
                mbar_arrive_expect_tx(v_full, k_chunks * SP_CHUNK_BYTES);
                for (int c = 0; c < k_chunks; ++c)
                    tma_load_3d(s_v + c * SP_CHUNK_BYTES, &tm_qkv, v_full, 2 * inner + h * SA_DH, c * 128, bf);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // One elected thread; descriptors are constants plus an address field (the GEMM's lesson, profiles/r1f:
        // an `if (lane == 0)` region costs ~26 SASS instructions per tcgen05.mma, an elect.sync region ~2).
        if (elect_one()) {
            const uint32_t idesc_s = make_idesc_bf16(SA_BM, 128, 0, 0);
            const uint32_t idesc_pv = make_idesc_bf16(SA_BM, SA_DH, 0, 1);   // B (= V) is MN-major
            const uint64_t desc_kmaj = make_smem_desc(0, 0, 1024, SWZ_128B);
            const uint64_t desc_v = make_smem_desc(smem_u32(s_v), 64 * 128, 1024, SWZ_128B);
            const uint32_t q_field = (smem_u32(s_q) & 0x3FFFFu) >> 4;
            const uint32_t k_field = (smem_u32(s_k) & 0x3FFFFu) >> 4;
            const int last_ksteps = (tokens - (k_chunks - 1) * 128 + 15) / 16;
            int qt = 0, n = 0;          // tile t = n * q_tiles + qt
            int qp = 0, np = 0;         // tile t - 1
            for (int t = 0; t <= n_tiles; ++t) {
                const int slot = t & 3;
                const int kb = n & 1;
                const int tp = t - 1;
                const int ob = tp & 1;
                const uint64_t q_desc = desc_kmaj | (q_field + slot * (SP_CHUNK_BYTES >> 4));
                const uint64_t k_desc = desc_kmaj | (k_field + kb * (SA_KV_BYTES >> 4));
                const uint32_t d_o = tmem_o + ob * SA_DH;
                for (int c = 0; c < k_chunks; ++c) {
                    if (t > 0) {
                        // ---- O(t-1) += P(t-1).c V.c ----  (P: 16 keys = 8 packed columns; part q owns columns 32q..)
                        if (c == 0) {
                            if (qp == 0) mbar_wait(v_full, np & 1);
                            mbar_wait(o_empty + ob, ((tp >> 1) & 1) ^ 1);
                        }
                        mbar_wait_hot(p_full + c, tp & 1);
                        tc_fence_after();
                        const uint32_t a0 = tmem_s + c * 128;
                        const uint64_t b0 = desc_v + static_cast<uint64_t>(c * 8 * (16 * 128 >> 4));
                        if (c != k_chunks - 1) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                umma_f16_ts(d_o, a0 + (j >> 1) * 32 + (j & 1) * 8, b0 + j * (16 * 128 >> 4), idesc_pv,
                                            (c | j) != 0 ? 1u : 0u);
                        } else {
                            for (int j = 0; j < last_ksteps; ++j)
                                umma_f16_ts(d_o, a0 + (j >> 1) * 32 + (j & 1) * 8, b0 + j * (16 * 128 >> 4), idesc_pv,
                                            (c | j) != 0 ? 1u : 0u);
                            umma_commit(o_full + ob);
                            if (qp == q_tiles - 1) umma_commit(v_empty);
                        }
                    }
                    if (t < n_tiles) {
                        // ---- S(t).c = Q(t) K.c^T ----
                        if (c == 0) {
                            mbar_wait_hot(q_full + slot, (t >> 2) & 1);
                            if (qt == 0) mbar_wait_hot(k_full + kb, (n >> 1) & 1);
                            tc_fence_after();
                        }
                        const uint64_t kc = k_desc + static_cast<uint64_t>(c * (SP_CHUNK_BYTES >> 4));
                        umma_f16_ss(tmem_s + c * 128, q_desc, kc, idesc_s, 0u);
                        umma_f16_ss(tmem_s + c * 128, q_desc + 2, kc + 2, idesc_s, 1u);
                        umma_f16_ss(tmem_s + c * 128, q_desc + 4, kc + 4, idesc_s, 1u);
                        umma_f16_ss(tmem_s + c * 128, q_desc + 6, kc + 6, idesc_s, 1u);
                        umma_commit(s_full + c);
                        if (c == k_chunks - 1) {
                            umma_commit(q_empty + slot);
                            if (qt == q_tiles - 1) umma_commit(k_empty + kb);
                        }
                    }
                }
                qp = qt; np = n;
                if (++qt == q_tiles) { qt = 0; ++n; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ================= softmax + epilogue =================
        const int quad = warp & 3;
        const int part = (warp - 4) >> 2;                 // which 32 columns of every 128-key chunk
        const int row = quad * 32 + lane;                 // row inside the q tile == TMEM lane
        const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t t_row = tmem_s + lane_base + part * 32;   // this thread's 32 columns of chunk 0
        float inv_prev = 0.0f;
        int64_t out_prev = -1;                             // element offset of this thread's 16 outputs, -1 = no store

        auto epilogue = [&](int tp) {
            const int ob = tp & 1;
            mbar_wait(o_full + ob, (tp >> 1) & 1);
            tc_fence_after();
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem_o + ob * SA_DH + lane_base + part * 16, r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty + ob);
            if (out_prev >= 0) {
                __nv_bfloat16* op = out + out_prev;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(r[8 * g + 0]) * inv_prev, __uint_as_float(r[8 * g + 1]) * inv_prev);
                    o.y = pack_bf16x2(__uint_as_float(r[8 * g + 2]) * inv_prev, __uint_as_float(r[8 * g + 3]) * inv_prev);
                    o.z = pack_bf16x2(__uint_as_float(r[8 * g + 4]) * inv_prev, __uint_as_float(r[8 * g + 5]) * inv_prev);
                    o.w = pack_bf16x2(__uint_as_float(r[8 * g + 6]) * inv_prev, __uint_as_float(r[8 * g + 7]) * inv_prev);
                    *reinterpret_cast<uint4*>(op + 8 * g) = o;
                }
            }
        };

        int qt = 0, n = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int item = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
            const int h = item % heads;
            const int bf = item / heads;
            const int q_idx = qt * SA_BM + row;

            // ---- pass 1: row max over the valid keys ----
            float mx = -INFINITY;
            for (int c = 0; c < k_chunks; ++c) {
                mbar_wait(s_full + c, t & 1);
                tc_fence_after();
                const int key0 = c * 128 + part * 32;
                if (key0 >= tokens) continue;              // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_row + c * 128, r);
                tmem_ld_wait();
                if (key0 + 32 <= tokens) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (key0 + j < tokens) mx = fmaxf(mx, __uint_as_float(r[j]));
                }
            }
            s_max[part * 128 + row] = mx;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * SP_SM_WARPS) : "memory");
            mx = fmaxf(fmaxf(s_max[row], s_max[128 + row]), fmaxf(s_max[256 + row], s_max[384 + row]));
            const float mxs = mx * scale_log2;

            // ---- deferred epilogue of the previous tile (its PV retired while pass 1 ran) ----
            if (t > 0) epilogue(t - 1);

            // ---- pass 2: P = exp2(S*c - max*c) -> bf16 -> TMEM (over the S columns just read), row sums ----
            // Instruction budget per element (profiles/r1m: the first version spent 7.5, XU and ALU pipes both
            // saturated): FFMA + MUFU.EX2 + FADD + integer rounding to bf16 (F2FP sits on the XU pipe with the
            // exponentials).  The denominator sums the unrounded exponentials.
            float sum = 0.0f;
            for (int c = 0; c < k_chunks; ++c) {
                uint32_t pk[16];
                const int key0 = c * 128 + part * 32;
                if (key0 >= tokens) {                      // warp-uniform: fully masked
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = 0u;
                } else {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_row + c * 128, r);
                    tmem_ld_wait();
                    if (key0 + 32 <= tokens) {             // warp-uniform: no masking
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float e0 = ex2_approx(fmaf(__uint_as_float(r[2 * j]), scale_log2, -mxs));
                            const float e1 = ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), scale_log2, -mxs));
                            sum += e0 + e1;
                            pk[j] = pack_bf16x2_rne_alu(e0, e1);
                        }
                    } else {
                        const int nvalid = tokens - key0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float e0 = ex2_approx(fmaf(__uint_as_float(r[2 * j]), scale_log2, -mxs));
                            float e1 = ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), scale_log2, -mxs));
                            if (2 * j >= nvalid) e0 = 0.0f;
                            if (2 * j + 1 >= nvalid) e1 = 0.0f;
                            sum += e0 + e1;
                            pk[j] = pack_bf16x2_rne_alu(e0, e1);
                        }
                    }
                }
                tmem_st_32x32b_x16(t_row + c * 128, pk);   // keys [32*part, 32*part+32) of chunk c -> 16 columns
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full + c);
            }
            s_sum[part * 128 + row] = sum;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * SP_SM_WARPS) : "memory");
            const float total = (s_sum[row] + s_sum[128 + row]) + (s_sum[256 + row] + s_sum[384 + row]);
            inv_prev = 1.0f / total;
            out_prev = (q_idx < tokens)
                           ? (static_cast<int64_t>(bf) * tokens + q_idx) * inner + h * SA_DH + part * 16
                           : -1;
            if (lse != nullptr && part == 0 && q_idx < tokens)
                lse[(static_cast<int64_t>(bf) * heads + h) * tokens + q_idx] = mxs + log2f(total);
            if (++qt == q_tiles) { qt = 0; ++n; }
        }
        if (n_tiles > 0) epilogue(n_tiles - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, SA_TMEM_COLS);
    }
}

// Two restructurings of the softmax side of this kernel were built, validated and measured in round 3 and did NOT
// pay (profiles/README.md r3c / r3d; the code is in the git history):
//   * two softmax groups alternating query tiles (ping-pong, barrier sets per tile parity): 0.289 vs 0.288 ms per
//     launch at the C2 size — the softmax warps' serial max -> barrier -> exp chain is not the limiter;
//   * S read from TMEM once (online row max per 128-key chunk, lazy O rescale, quad-level named barriers):
//     0.338 ms — the per-chunk barrier puts the 4 warps of a scheduler in lockstep, so tcgen05.ld latency no longer
//     overlaps the MUFU work of sibling warps; halving the TMEM reads (64 B/clk port) did not compensate.

}  // namespace istvt
#include "attn_spatial_pp.cuh"
#include "attn_spatial_pp3.cuh"
namespace istvt {

// ------------------------------------------------------------------------------------------
// fp32 validation kernel: one CTA per (frame, head); K and V in shared memory, one query per thread.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_spatial_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ probs,
                        int tokens, int heads, float scale) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* sk = reinterpret_cast<float*>(smem_raw);  // [tokens][64]
    float* sv = sk + static_cast<size_t>(tokens) * SA_DH;
    const int h = blockIdx.x % heads;
    const int bf = blockIdx.x / heads;
    const int inner = heads * SA_DH;
    const int64_t row0 = static_cast<int64_t>(bf) * tokens;
    for (int c = threadIdx.x; c < tokens * (SA_DH / 4); c += blockDim.x) {
        const int j = c / (SA_DH / 4);
        const int d = (c - j * (SA_DH / 4)) * 4;
        const float* base = qkv + (row0 + j) * (3 * inner) + h * SA_DH + d;
        *reinterpret_cast<float4*>(sk + j * SA_DH + d) = *reinterpret_cast<const float4*>(base + inner);
        *reinterpret_cast<float4*>(sv + j * SA_DH + d) = *reinterpret_cast<const float4*>(base + 2 * inner);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tokens; i += blockDim.x) {
        float q[SA_DH], o[SA_DH];
        const float* qp = qkv + (row0 + i) * (3 * inner) + h * SA_DH;
#pragma unroll
        for (int d = 0; d < SA_DH; d += 4) {
            const float4 t = *reinterpret_cast<const float4*>(qp + d);
            q[d] = t.x * scale; q[d + 1] = t.y * scale; q[d + 2] = t.z * scale; q[d + 3] = t.w * scale;
        }
#pragma unroll
        for (int d = 0; d < SA_DH; ++d) o[d] = 0.0f;
        float mx = -INFINITY, l = 0.0f;
        for (int j = 0; j < tokens; ++j) {
            float s = 0.0f;
#pragma unroll
            for (int d = 0; d < SA_DH; d += 4) {
                const float4 kk = *reinterpret_cast<const float4*>(sk + j * SA_DH + d);
                s = fmaf(q[d], kk.x, s); s = fmaf(q[d + 1], kk.y, s);
                s = fmaf(q[d + 2], kk.z, s); s = fmaf(q[d + 3], kk.w, s);
            }
            const float mnew = fmaxf(mx, s);
            const float corr = expf(mx - mnew);
            const float pj = expf(s - mnew);
            l = l * corr + pj;
#pragma unroll
            for (int d = 0; d < SA_DH; d += 4) {
                const float4 vv = *reinterpret_cast<const float4*>(sv + j * SA_DH + d);
                o[d] = fmaf(o[d], corr, pj * vv.x); o[d + 1] = fmaf(o[d + 1], corr, pj * vv.y);
                o[d + 2] = fmaf(o[d + 2], corr, pj * vv.z); o[d + 3] = fmaf(o[d + 3], corr, pj * vv.w);
            }
            mx = mnew;
        }
        const float inv = 1.0f / l;
        float* op = out + (row0 + i) * inner + h * SA_DH;
#pragma unroll
        for (int d = 0; d < SA_DH; d += 4)
            *reinterpret_cast<float4*>(op + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
        if (probs != nullptr) {
            float* pr = probs + ((static_cast<int64_t>(bf) * heads + h) * tokens + i) * tokens;
            for (int j = 0; j < tokens; ++j) {
                float s = 0.0f;
#pragma unroll
                for (int d = 0; d < SA_DH; d += 4) {
                    const float4 kk = *reinterpret_cast<const float4*>(sk + j * SA_DH + d);
                    s = fmaf(q[d], kk.x, s); s = fmaf(q[d + 1], kk.y, s);
                    s = fmaf(q[d + 2], kk.z, s); s = fmaf(q[d + 3], kk.w, s);
                }
                pr[j] = expf(s - mx) * inv;
            }
        }
    }
}

}  // namespace istvt

using namespace istvt;

static int attn_spatial_launch(const void* qkv, void* out, float* probs, float* lse, int dtype, int batch_frames,
                               int tokens, int heads, float scale, cudaStream_t st) {
    ISTVT_REQUIRE(qkv && out);
    ISTVT_REQUIRE(batch_frames > 0 && tokens > 0 && heads > 0);
    const int inner = heads * SA_DH;
    if (dtype == ISTVT_F32) {
        ISTVT_REQUIRE(lse == nullptr);
        const size_t smem = 2 * static_cast<size_t>(tokens) * SA_DH * sizeof(float);
        ISTVT_REQUIRE(smem <= 220 * 1024);
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(smem)));
        attn_spatial_f32_kernel<<<batch_frames * heads, 128, smem, st>>>(
            static_cast<const float*>(qkv), static_cast<float*>(out), probs, tokens, heads, scale);
        count_launch();
        return launch_status();
    }
    ISTVT_REQUIRE(dtype == ISTVT_BF16);
    ISTVT_REQUIRE(tokens <= SA_KMAX);
    ISTVT_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
    CUtensorMap tm;
    {
        const uint64_t dims[3] = {static_cast<uint64_t>(3 * inner), static_cast<uint64_t>(tokens),
                                  static_cast<uint64_t>(batch_frames)};
        const uint64_t strides[2] = {static_cast<uint64_t>(3 * inner) * 2,
                                     static_cast<uint64_t>(tokens) * 3 * inner * 2};
        const uint32_t box[3] = {SA_DH, 128, 1};
        int rc = encode_tmap(&tm, qkv, ISTVT_BF16, 3, dims, strides, box, 3);
        if (rc != ISTVT_OK) return rc;
    }
    const float scale_log2 = scale * 1.4426950408889634f;
    if (probs == nullptr) {
        // production path: persistent kernel, one CTA per SM.  Inference default: two query tiles in flight, thread = query
        // row, online softmax (attn_spatial_pp.cuh).  The training / relevance forward (lse wanted) stays on the exact-max
        // kernel below: there the dominant key of a row is exactly 1.0 in bf16, and the spatial relevance maps — products
        // of gradients and probabilities through 12 layers — react to that last bit (profiles/README.md r7o / r7p).
        // ISTVT_SA_KERNEL=pipe | pp | pp3 forces one kernel for A/B runs (pp3: inference only).
        const int items = batch_frames * heads;
        const int grid = items < sm_count() ? items : sm_count();
        const char* sel = getenv("ISTVT_SA_KERNEL");
        const bool use_pp = sel != nullptr ? strcmp(sel, "pipe") != 0 : lse == nullptr;
        const bool use_pp3 = sel != nullptr && strcmp(sel, "pp3") == 0 && lse == nullptr &&
                             static_cast<int64_t>(batch_frames) * tokens * inner < (int64_t(1) << 31);
        if (use_pp3) {
            // three query tiles in flight (attn_spatial_pp3.cuh): measured equal to the two-tile kernel, opt-in
            ISTVT_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_pp3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  S3_SMEM));
            attn_spatial_pp3_kernel<<<grid, S3_THREADS, S3_SMEM, st>>>(tm, static_cast<__nv_bfloat16*>(out), tokens, heads,
                                                                      items, scale_log2);
            count_launch();
            return launch_status();
        }
        if (use_pp) {
            const char* rnd = getenv("ISTVT_SA_ROUND");
            auto kern = (rnd != nullptr && strcmp(rnd, "trunc") == 0) ? attn_spatial_pp_kernel<false> : attn_spatial_pp_kernel<true>;
            ISTVT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM));
            kern<<<grid, S2_THREADS, S2_SMEM, st>>>(tm, static_cast<__nv_bfloat16*>(out), lse, tokens, heads, items,
                                                    scale_log2);
            count_launch();
            return launch_status();
        }
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              SP_SMEM));
        attn_spatial_pipe_kernel<<<grid, SP_THREADS, SP_SMEM, st>>>(tm, static_cast<__nv_bfloat16*>(out), lse, tokens,
                                                                   heads, items, scale_log2);
        count_launch();
        return launch_status();
    }
    // attention-map mode (parity tests, relevance pass): one CTA per (frame, head, query tile), S kept in TMEM
    ISTVT_REQUIRE(lse == nullptr);
    ISTVT_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          SA_SMEM));
    const int q_tiles = (tokens + SA_BM - 1) / SA_BM;
    attn_spatial_tcgen05_kernel<<<batch_frames * heads * q_tiles, SA_THREADS, SA_SMEM, st>>>(
        tm, static_cast<__nv_bfloat16*>(out), probs, tokens, heads, scale_log2);
    count_launch();
    return launch_status();
}

extern "C" int istvt_attn_spatial_fwd(const void* qkv, void* out, float* probs, int dtype, int batch_frames,
                                      int tokens, int heads, float scale, istvt_stream_t stream) {
    return attn_spatial_launch(qkv, out, probs, nullptr, dtype, batch_frames, tokens, heads, scale,
                               static_cast<cudaStream_t>(stream));
}

// Training-mode forward: same kernel, additionally writes lse[batch_frames, heads, tokens] (fp32, log2 domain:
// log2(sum_j exp2(s_ij * scale * log2 e))) for istvt_attn_spatial_bwd.  bf16 only.
extern "C" int istvt_attn_spatial_fwd_lse(const void* qkv, void* out, float* lse, int batch_frames, int tokens,
                                          int heads, float scale, istvt_stream_t stream) {
    ISTVT_REQUIRE(lse != nullptr);
    return attn_spatial_launch(qkv, out, nullptr, lse, ISTVT_BF16, batch_frames, tokens, heads, scale,
                               static_cast<cudaStream_t>(stream));
}
