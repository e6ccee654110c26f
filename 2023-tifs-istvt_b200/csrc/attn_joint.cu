// Joint self-attention over ALL tokens of a sequence, any sequence length (network/vivit/module.py:53-63,
// `Attention.forward`: the ablation transformers `Transformer` / `ViViT` / `VanillaTr`, vivit.py:10-25,29-81,150-191).
// At VanillaTr's 6*361+1 = 2167 tokens the score matrix no longer fits TMEM (the spatial kernel keeps all 384
// key columns resident), so these kernels stream the keys in blocks of 128 with an online softmax.
//
// Two bf16 kernels (tcgen05).  Sequences of at least two query tiles run attn_joint_pp_kernel (attn_joint_pp.cuh: two
// tiles per CTA in ping-pong, the faster one); this file holds the one-tile-per-CTA kernel it grew out of, which serves
// single-tile sequences: one CTA = one (sequence b, head h, 128-query tile).
//   warp 0     TMA producer: Q once, then K / V blocks of 128 keys through a 4-stage ring, read in place from the
//              packed projection output [rows, 3*heads*64] (3-D tensor map: col, token, sequence; OOB rows zero-filled)
//   warp 1     MMA issuer:   S_j[128 x 128] = Q K_j^T      (both operands K-major SW128, accumulator in TMEM)
//                            O_j[128 x 64]  = P_j V_j       (P_j bf16 in smem, V_j the MN-major B operand)
//              S_{j+1} is issued before the issuer waits for P_j, so the next score block is ready when the softmax
//              warps come back for it
//   warps 4-11 softmax: 2 threads per query row (64 key columns of the block each): running maximum m, running
//              denominator l, P_j = exp2(S_j c - m c) -> smem; the per-block product O_j is added into a REGISTER
//              accumulator o = o * exp2((m_old - m_new) c) + O_j one block late (while S_{j+1} / PV_j run), so no
//              accumulator in TMEM ever needs rescaling.  The two threads of a row sit in warps quad and quad + 4 and
//              exchange their block maxima through a named barrier of their own (no CTA-wide lockstep); P is rounded
//              to bf16 on the integer pipe.
//   TMEM: S0 S1 [128 x 128] fp32 (cols 0-255), O0 O1 [128 x 64] fp32 (cols 256-383).
//   smem: Q 16 KB, K 4 x 16 KB, V 4 x 16 KB (ring of 4 key blocks), P 2 x 32 KB.
// fp32 path: SIMT validation kernel, one query per thread, K / V streamed through shared memory in blocks of 64.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "simt_util.cuh"

namespace istvt {

constexpr int JA_DH = 64;
constexpr int JA_BM = 128;         // queries per CTA
constexpr int JA_BN = 128;         // keys per block
// warps 0-3: TMA / MMA / TMEM alloc / spare; then SPLIT x 4 softmax warps (SPLIT threads per query row)
constexpr int ja_threads(int split) { return 128 + split * 128; }
constexpr int JA_TILE_BYTES = 128 * JA_DH * 2;       // 16 KB: a Q tile, a K block or a V block
constexpr int JA_P_BYTES = JA_BM * JA_BN * 2;        // 32 KB: two 64-key SW128 atoms of 128 rows
constexpr int JA_STAGES = 4;       // K / V ring: the fetch of block j+3 starts when PV_{j-1} retires (a 2-stage ring left
                                   // the TMA latency of K_{j+1} exposed: the softmax warps' top stall was the wait for S)
constexpr int JA_SMEM = (1 + 2 * JA_STAGES) * JA_TILE_BYTES + 2 * JA_P_BYTES + 8192 + 1024;
constexpr int JA_TMEM_COLS = 512;

__device__ __forceinline__ void tmem_ld_32x32b(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32b_x32(taddr, r); }
__device__ __forceinline__ void tmem_ld_32x32b(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x32b_x16(taddr, r); }

// row maximum of this thread's slice of a score block (MASK: columns >= valid are past the sequence's last key)
template <int NLD, bool MASK>
__device__ __forceinline__ float block_max(const uint32_t (&r)[NLD][32], int valid) {
    float mx = -INFINITY;
#pragma unroll
    for (int l = 0; l < NLD; ++l)
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (!MASK || l * 32 + c < valid) mx = fmaxf(mx, __uint_as_float(r[l][c]));
    return mx;
}

// p = exp2(s * c - m * c) for this thread's slice, rounded to bf16 and stored as the PV MMA's A operand; returns the
// slice's contribution to the row denominator, accumulated from the SAME rounded values the MMA consumes.
template <int NLD, bool MASK, bool ALUPACK, bool ROUNDED_SUM = true>
__device__ __forceinline__ float block_exp_store(const uint32_t (&r)[NLD][32], int valid, float scale_log2, float mxs,
                                                 uint8_t* prow, int chunk0, int row) {
    float sum = 0.0f;
#pragma unroll
    for (int l = 0; l < NLD; ++l) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int c = l * 32 + 2 * i;
            float e0 = ex2_approx(fmaf(__uint_as_float(r[l][2 * i]), scale_log2, -mxs));
            float e1 = ex2_approx(fmaf(__uint_as_float(r[l][2 * i + 1]), scale_log2, -mxs));
            if (MASK) {
                e0 = (c < valid) ? e0 : 0.0f;
                e1 = (c + 1 < valid) ? e1 : 0.0f;
            }
            if (ALUPACK) {   // round on the integer pipe: F2FP shares the XU pipe with the exponentials
                pk[i] = pack_bf16x2_rne_alu(e0, e1);
                if (ROUNDED_SUM) sum += __uint_as_float(pk[i] << 16) + __uint_as_float(pk[i] & 0xffff0000u);
                else sum += e0 + e1;   // round-to-nearest is unbiased: the spatial kernel sums the unrounded values too
            } else {
                pk[i] = pack_bf16x2(e0, e1);
                const float2 f = unpack_bf16x2(pk[i]);
                sum += f.x + f.y;
            }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int chunk = (chunk0 + l * 4 + g) ^ (row & 7);
            *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
    }
    return sum;
}

template <int SPLIT, bool QUADBAR, bool ALUPACK>
__global__ void __launch_bounds__(ja_threads(SPLIT), 1)
attn_joint_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, int tokens,
                          int heads, float scale_log2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_q = smem;
    uint8_t* s_k = s_q + JA_TILE_BYTES;                    // [JA_STAGES][16 KB]
    uint8_t* s_v = s_k + JA_STAGES * JA_TILE_BYTES;        // [JA_STAGES][16 KB]
    uint8_t* s_p = s_v + JA_STAGES * JA_TILE_BYTES;        // [2][32 KB]
    uint8_t* s_misc = s_p + 2 * JA_P_BYTES;
    uint64_t* bar_q = reinterpret_cast<uint64_t*>(s_misc);
    uint64_t* kv_full = bar_q + 1;                         // [JA_STAGES]
    uint64_t* kv_empty = kv_full + JA_STAGES;              // [JA_STAGES]
    uint64_t* bar_s = kv_empty + JA_STAGES;                // [2]
    uint64_t* bar_p = bar_s + 2;                           // [2]
    uint64_t* bar_o = bar_p + 2;                           // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_o + 2);
    float* s_red = reinterpret_cast<float*>(s_misc + 256);   // [2 block parities][SPLIT slices][128 rows], then sums

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int q_tiles = (tokens + JA_BM - 1) / JA_BM;
    const int qt = blockIdx.x % q_tiles;
    const int h = (blockIdx.x / q_tiles) % heads;
    const int b = blockIdx.x / (q_tiles * heads);
    const int inner = heads * JA_DH;
    const int nblk = (tokens + JA_BN - 1) / JA_BN;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
    if (warp == 1 && lane == 0) {
        mbar_init(bar_q, 1);
        for (int s = 0; s < JA_STAGES; ++s) {
            mbar_init(kv_full + s, 1);
            mbar_init(kv_empty + s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_s + s, 1);
            mbar_init(bar_p + s, SPLIT * 4);   // one arrive per softmax warp
            mbar_init(bar_o + s, 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_holder, JA_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tmem_s = tmem_base;              // + s * 128
    const uint32_t tmem_o = tmem_base + 2 * JA_BN;  // + s * 64

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_q, JA_TILE_BYTES);
            tma_load_3d(s_q, &tm_qkv, bar_q, h * JA_DH, qt * JA_BM, b);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % JA_STAGES;
                if (j >= JA_STAGES) mbar_wait_sleep(kv_empty + s, ((j / JA_STAGES) - 1) & 1);
                mbar_arrive_expect_tx(kv_full + s, 2 * JA_TILE_BYTES);
                tma_load_3d(s_k + s * JA_TILE_BYTES, &tm_qkv, kv_full + s, inner + h * JA_DH, j * JA_BN, b);
                tma_load_3d(s_v + s * JA_TILE_BYTES, &tm_qkv, kv_full + s, 2 * inner + h * JA_DH, j * JA_BN, b);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        const uint32_t idesc_s = make_idesc_bf16(JA_BM, JA_BN, 0, 0);
        const uint32_t idesc_o = make_idesc_bf16(JA_BM, JA_DH, 0, 1);   // B (= V) is MN-major
        const uint32_t q_addr = smem_u32(s_q);
        mbar_wait(bar_q, 0);
        mbar_wait(kv_full + 0, 0);
        tc_fence_after();
        if (lane == 0) {
            const uint32_t k_addr = smem_u32(s_k);
#pragma unroll
            for (int k = 0; k < JA_DH / 16; ++k)
                umma_f16_ss(tmem_s, make_smem_desc(q_addr + k * 32, 0, 1024, SWZ_128B),
                            make_smem_desc(k_addr + k * 32, 0, 1024, SWZ_128B), idesc_s, k != 0 ? 1u : 0u);
            umma_commit(bar_s + 0);
        }
        __syncwarp();
        for (int j = 0; j < nblk; ++j) {
            if (j + 1 < nblk) {   // S_{j+1}: its TMEM buffer was released by bar_p of block j-1 (waited last iteration)
                const int s1 = (j + 1) & 1;
                const int st1 = (j + 1) % JA_STAGES;
                mbar_wait(kv_full + st1, ((j + 1) / JA_STAGES) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t k_addr = smem_u32(s_k + st1 * JA_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < JA_DH / 16; ++k)
                        umma_f16_ss(tmem_s + s1 * JA_BN, make_smem_desc(q_addr + k * 32, 0, 1024, SWZ_128B),
                                    make_smem_desc(k_addr + k * 32, 0, 1024, SWZ_128B), idesc_s, k != 0 ? 1u : 0u);
                    umma_commit(bar_s + s1);
                }
                __syncwarp();
            }
            const int s = j & 1;
            mbar_wait(bar_p + s, (j >> 1) & 1);   // P_j in smem; O buffer s drained (block j-2 was added before P_j)
            tc_fence_after();
            if (lane == 0) {
                const uint32_t p_addr = smem_u32(s_p + s * JA_P_BYTES);
                const uint32_t v_addr = smem_u32(s_v + (j % JA_STAGES) * JA_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < JA_BN / 16; ++k) {
                    const uint64_t a_desc =
                        make_smem_desc(p_addr + (k >> 2) * (JA_BM * 128) + (k & 3) * 32, 0, 1024, SWZ_128B);
                    const uint64_t b_desc = make_smem_desc(v_addr + k * 16 * 128, 64 * 128, 1024, SWZ_128B);
                    umma_f16_ss(tmem_o + s * JA_DH, a_desc, b_desc, idesc_o, k != 0 ? 1u : 0u);
                }
                umma_commit(bar_o + s);
                umma_commit(kv_empty + (j % JA_STAGES));   // K_j (read by S_j) and V_j are free once PV_j retires
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        constexpr int COLS = JA_BN / SPLIT;             // key columns of a block per thread: 64 | 32
        constexpr int NLD = COLS / 32;                  // 32-column TMEM loads per block: 2 | 1
        constexpr int OC = JA_DH / SPLIT;               // output columns per thread: 32 | 16
        const int quad = warp & 3;
        const int part = (warp - 4) >> 2;               // which column slice of the row this thread owns
        const int row = quad * 32 + lane;               // row inside the q tile == TMEM lane
        const int q_idx = qt * JA_BM + row;
        const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;

        float m_run = -INFINITY, l_part = 0.0f, corr_prev = 0.0f;
        float o[OC];
#pragma unroll
        for (int i = 0; i < OC; ++i) o[i] = 0.0f;

        auto add_block = [&](int sp) {                  // o = o * corr_prev + O_sp
            uint32_t ro[OC];
            tmem_ld_32x32b(tmem_o + lane_base + sp * JA_DH + part * OC, ro);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < OC; ++i) o[i] = fmaf(o[i], corr_prev, __uint_as_float(ro[i]));
        };

        for (int j = 0; j < nblk; ++j) {
            const int s = j & 1;
            mbar_wait(bar_s + s, (j >> 1) & 1);
            tc_fence_after();
            uint32_t r[NLD][32];
            const uint32_t t_s = tmem_s + lane_base + s * JA_BN + part * COLS;
#pragma unroll
            for (int l = 0; l < NLD; ++l) tmem_ld_32x32b_x32(t_s + l * 32, r[l]);
            tmem_ld_wait();
            const int valid = tokens - (j * JA_BN + part * COLS);   // columns [0, valid) of this slice are real keys
            // only the sequence's last key block has columns past the last token: everything else skips the masks
            const bool ragged = valid < COLS;            // warp-uniform
            const float mx = ragged ? block_max<NLD, true>(r, valid) : block_max<NLD, false>(r, valid);
            float* red = s_red + s * (SPLIT * 128);
            red[part * 128 + row] = mx;
            // the row's SPLIT threads sit in warps quad, quad + 4, ...: QUADBAR exchanges inside that group only
            // (named barrier 1 + quad, SPLIT warps) instead of across all softmax warps
            if (QUADBAR) asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(SPLIT * 32) : "memory");
            else asm volatile("bar.sync 1, %0;" ::"n"(SPLIT * 128) : "memory");
            float m_new = m_run;
#pragma unroll
            for (int p = 0; p < SPLIT; ++p) m_new = fmaxf(m_new, red[p * 128 + row]);
            const float corr = ex2_approx((m_run - m_new) * scale_log2);
            const float mxs = m_new * scale_log2;

            // P_j, bf16, K-major SW128: atom = 64 keys (128 B per row), 16-byte chunks XOR-swizzled by row & 7
            uint8_t* prow = s_p + s * JA_P_BYTES + ((part * COLS) >> 6) * (JA_BM * 128) + (row >> 3) * 1024 +
                            (row & 7) * 128;
            const int chunk0 = ((part * COLS) & 63) >> 3;
            const float sum = ragged ? block_exp_store<NLD, true, ALUPACK>(r, valid, scale_log2, mxs, prow, chunk0, row)
                                     : block_exp_store<NLD, false, ALUPACK>(r, valid, scale_log2, mxs, prow, chunk0, row);
            l_part = fmaf(l_part, corr, sum);
            fence_proxy_async_smem();   // st.shared P -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p + s);

            if (j > 0) {   // deferred: o = o * corr_{j-1} + O_{j-1}
                mbar_wait(bar_o + ((j - 1) & 1), ((j - 1) >> 1) & 1);
                tc_fence_after();
                add_block((j - 1) & 1);
            }
            corr_prev = corr;
            m_run = m_new;
        }
        mbar_wait(bar_o + ((nblk - 1) & 1), ((nblk - 1) >> 1) & 1);
        tc_fence_after();
        add_block((nblk - 1) & 1);
        // denominators of the column slices (each already in the scale of the final maximum)
        float* sums = s_red + 2 * SPLIT * 128;
        sums[part * 128 + row] = l_part;
        if (QUADBAR) asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(SPLIT * 32) : "memory");
        else asm volatile("bar.sync 1, %0;" ::"n"(SPLIT * 128) : "memory");
        float total = 0.0f;
#pragma unroll
        for (int p = 0; p < SPLIT; ++p) total += sums[p * 128 + row];
        const float inv = 1.0f / total;
        if (q_idx < tokens) {
            __nv_bfloat16* op = out + (static_cast<int64_t>(b) * tokens + q_idx) * inner + h * JA_DH + part * OC;
#pragma unroll
            for (int g = 0; g < OC / 8; ++g) {
                uint4 v;
                v.x = pack_bf16x2(o[8 * g + 0] * inv, o[8 * g + 1] * inv);
                v.y = pack_bf16x2(o[8 * g + 2] * inv, o[8 * g + 3] * inv);
                v.z = pack_bf16x2(o[8 * g + 4] * inv, o[8 * g + 5] * inv);
                v.w = pack_bf16x2(o[8 * g + 6] * inv, o[8 * g + 7] * inv);
                *reinterpret_cast<uint4*>(op + 8 * g) = v;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, JA_TMEM_COLS);
    }
}

}  // namespace istvt

#include "attn_joint_pp.cuh"

namespace istvt {

// ------------------------------------------------------------------------------------------
// fp32 validation kernel: one CTA per (sequence, head, 128 queries); K and V stream through shared memory in
// blocks of 64 keys, one query per thread with an online softmax.
// ------------------------------------------------------------------------------------------
constexpr int JF_KB = 64;

__global__ void __launch_bounds__(128)
attn_joint_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int tokens, int heads, float scale) {
    __shared__ __align__(16) float sk[JF_KB * JA_DH];
    __shared__ __align__(16) float sv[JF_KB * JA_DH];
    const int q_tiles = (tokens + 127) / 128;
    const int qt = blockIdx.x % q_tiles;
    const int h = (blockIdx.x / q_tiles) % heads;
    const int b = blockIdx.x / (q_tiles * heads);
    const int inner = heads * JA_DH;
    const int64_t row0 = static_cast<int64_t>(b) * tokens;
    const int i = qt * 128 + threadIdx.x;
    const bool live = i < tokens;

    float q[JA_DH], o[JA_DH];
    {
        const float* qp = qkv + (row0 + (live ? i : 0)) * (3 * inner) + h * JA_DH;
#pragma unroll
        for (int d = 0; d < JA_DH; d += 4) {
            const float4 t = *reinterpret_cast<const float4*>(qp + d);
            q[d] = t.x * scale; q[d + 1] = t.y * scale; q[d + 2] = t.z * scale; q[d + 3] = t.w * scale;
        }
    }
#pragma unroll
    for (int d = 0; d < JA_DH; ++d) o[d] = 0.0f;
    float mx = -INFINITY, l = 0.0f;

    for (int k0 = 0; k0 < tokens; k0 += JF_KB) {
        const int nk = min(JF_KB, tokens - k0);
        __syncthreads();
        for (int c = threadIdx.x; c < nk * (JA_DH / 4); c += blockDim.x) {
            const int j = c / (JA_DH / 4);
            const int d = (c - j * (JA_DH / 4)) * 4;
            const float* base = qkv + (row0 + k0 + j) * (3 * inner) + h * JA_DH + d;
            *reinterpret_cast<float4*>(sk + j * JA_DH + d) = *reinterpret_cast<const float4*>(base + inner);
            *reinterpret_cast<float4*>(sv + j * JA_DH + d) = *reinterpret_cast<const float4*>(base + 2 * inner);
        }
        __syncthreads();
        for (int j = 0; j < nk; ++j) {
            float s = 0.0f;
#pragma unroll
            for (int d = 0; d < JA_DH; d += 4) {
                const float4 kk = *reinterpret_cast<const float4*>(sk + j * JA_DH + d);
                s = fmaf(q[d], kk.x, s); s = fmaf(q[d + 1], kk.y, s);
                s = fmaf(q[d + 2], kk.z, s); s = fmaf(q[d + 3], kk.w, s);
            }
            const float mnew = fmaxf(mx, s);
            const float corr = expf(mx - mnew);
            const float pj = expf(s - mnew);
            l = l * corr + pj;
#pragma unroll
            for (int d = 0; d < JA_DH; d += 4) {
                const float4 vv = *reinterpret_cast<const float4*>(sv + j * JA_DH + d);
                o[d] = fmaf(o[d], corr, pj * vv.x); o[d + 1] = fmaf(o[d + 1], corr, pj * vv.y);
                o[d + 2] = fmaf(o[d + 2], corr, pj * vv.z); o[d + 3] = fmaf(o[d + 3], corr, pj * vv.w);
            }
            mx = mnew;
        }
    }
    if (live) {
        const float inv = 1.0f / l;
        float* op = out + (row0 + i) * inner + h * JA_DH;
#pragma unroll
        for (int d = 0; d < JA_DH; d += 4)
            *reinterpret_cast<float4*>(op + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
    }
}

}  // namespace istvt

using namespace istvt;

// Kernel choice.  Sequences of at least two query tiles run the ping-pong kernel (attn_joint_pp.cuh: 0.339 ms at 16 x 2167
// tokens); single-tile sequences, or ISTVT_JA_PP=0 (A/B switch), the one-tile-per-CTA kernel above (0.365 ms).
// Alternatives of the one-tile kernel that were measured through its template parameters and are not instantiated
// (profiles/README.md r5b-r5d): 4 softmax threads per row (0.390 ms), CTA-wide instead of per-quad exchange of the row
// maximum (0.378), F2FP instead of integer-pipe rounding of P (0.368), neither (0.385); of the ping-pong kernel: the
// denominator summed from the bf16-rounded P (0.340 vs 0.339), paired TMEM loads (0.338, r5q).
static int joint_pingpong() {
    static const int v = [] {
        const char* e = getenv("ISTVT_JA_PP");
        return e != nullptr ? atoi(e) : 1;
    }();
    return v;
}

extern "C" int istvt_attn_joint_fwd(const void* qkv, void* out, int dtype, int batch, int tokens, int heads,
                                    float scale, istvt_stream_t stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ISTVT_REQUIRE(qkv && out);
    ISTVT_REQUIRE(batch > 0 && tokens > 0 && heads > 0 && scale > 0.0f);
    const int inner = heads * JA_DH;
    const int q_tiles = (tokens + 127) / 128;
    ISTVT_REQUIRE(static_cast<int64_t>(batch) * heads * q_tiles < (int64_t(1) << 31));
    if (dtype == ISTVT_F32) {
        attn_joint_f32_kernel<<<batch * heads * q_tiles, 128, 0, st>>>(
            static_cast<const float*>(qkv), static_cast<float*>(out), tokens, heads, scale);
        count_launch();
        return launch_status();
    }
    ISTVT_REQUIRE(dtype == ISTVT_BF16);
    ISTVT_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
    CUtensorMap tm;
    {
        const uint64_t dims[3] = {static_cast<uint64_t>(3 * inner), static_cast<uint64_t>(tokens),
                                  static_cast<uint64_t>(batch)};
        const uint64_t strides[2] = {static_cast<uint64_t>(3 * inner) * 2,
                                     static_cast<uint64_t>(tokens) * 3 * inner * 2};
        const uint32_t box[3] = {JA_DH, 128, 1};
        int rc = encode_tmap(&tm, qkv, ISTVT_BF16, 3, dims, strides, box, 3);
        if (rc != ISTVT_OK) return rc;
    }
    const float scale_log2 = scale * 1.4426950408889634f;
    const int grid = batch * heads * q_tiles;
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
    if (joint_pingpong() && q_tiles >= 2) {
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(attn_joint_pp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              PP_SMEM));
        attn_joint_pp_kernel<false><<<batch * heads * ((q_tiles + 1) / 2), PP_THREADS, PP_SMEM, st>>>(tm, o, tokens, heads,
                                                                                                     scale_log2);
    } else {
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(attn_joint_tcgen05_kernel<2, true, true>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, JA_SMEM));
        attn_joint_tcgen05_kernel<2, true, true><<<grid, ja_threads(2), JA_SMEM, st>>>(tm, o, tokens, heads, scale_log2);
    }
    count_launch();
    return launch_status();
}
