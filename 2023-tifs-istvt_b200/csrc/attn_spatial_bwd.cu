// Backward of the spatial self-attention (network/vivit/module.py:84-91), bf16, flash-attention-2 style on the
// warp-level tensor-core path (mma.sync m16n8k16 + ldmatrix):
//   one CTA = (frame, head, block of 128 keys); it keeps dK / dV [128 x 64] in registers and walks the query
//   blocks of 64 rows:  S = Q K^T,  P = exp2(S * c - lse),  dP = dO V^T,  D = rowsum(dO o O),
//                       dS = P o (dP - D) * scale,  dV += P^T dO,  dK += dS^T Q,  dQ += dS K
//   dQ is accumulated across the (up to 3) key blocks with red.global.add.f32 into an fp32 scratch buffer and
//   converted to bf16 by a second small kernel.
// P uses the log-sum-exp saved by the forward (istvt_attn_spatial_fwd_lse), so no second softmax pass is needed.
// (A tcgen05 version is the follow-up; this kernel is ~5 % of the training step.)
#include "common.cuh"
#include "ptx.cuh"
#include "simt_util.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>

namespace istvt {

typedef __nv_bfloat16 bf16;

constexpr int SB_DH = 64;
constexpr int SB_KB = 128;            // keys per CTA
constexpr int SB_QB = 64;             // queries per iteration
constexpr int SB_LD = SB_DH + 8;      // smem row pitch (elements) of the [*, 64] tiles: 144 B, conflict-free ldmatrix
constexpr int SB_LDP = SB_KB + 8;     // pitch of the P / dS tiles: 272 B
constexpr int SB_THREADS = 512;   // 16 warps: 4 per scheduler hide the ldmatrix -> mma latency (8 warps: 105 TFLOP/s)
constexpr int SB_SMEM = (2 * SB_KB * SB_LD + 2 * SB_QB * SB_LD + 2 * SB_QB * SB_LDP) * 2 + 2 * SB_QB * 4;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pkbf(float a, float b) {
    const __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&r);
}

__global__ void __launch_bounds__(SB_THREADS, 1)
attn_spatial_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ dout,
                        const float* __restrict__ lse, bf16* __restrict__ dqkv, float* __restrict__ dq_acc,
                        float* __restrict__ cam, int tokens, int heads, float scale) {
    extern __shared__ __align__(16) uint8_t sb_smem[];
    bf16* sK = reinterpret_cast<bf16*>(sb_smem);          // [128][72]
    bf16* sV = sK + SB_KB * SB_LD;                        // [128][72]
    bf16* sQ = sV + SB_KB * SB_LD;                        // [64][72]
    bf16* sdO = sQ + SB_QB * SB_LD;                       // [64][72]
    bf16* sP = sdO + SB_QB * SB_LD;                       // [64][136]
    bf16* sdS = sP + SB_QB * SB_LDP;                      // [64][136]
    float* sD = reinterpret_cast<float*>(sdS + SB_QB * SB_LDP);   // [64]
    float* sL = sD + SB_QB;                                        // [64]

    const int k_blocks = (tokens + SB_KB - 1) / SB_KB;
    const int kb = blockIdx.x % k_blocks;
    const int h = (blockIdx.x / k_blocks) % heads;
    const int bf = blockIdx.x / (k_blocks * heads);
    const int inner = heads * SB_DH;
    const int64_t row0 = static_cast<int64_t>(bf) * tokens;
    const int key0 = kb * SB_KB;
    const float scale_log2 = scale * 1.4426950408889634f;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int lmat = lane >> 3, lr = lane & 7;

    // ---- K / V block -> smem (rows past the frame end are zero) ----
    for (int idx = tid; idx < SB_KB * 8; idx += SB_THREADS) {
        const int r = idx >> 3, c = (idx & 7) * 8;
        uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
        if (key0 + r < tokens) {
            const bf16* base = qkv + (row0 + key0 + r) * (3 * inner) + h * SB_DH + c;
            kv = *reinterpret_cast<const uint4*>(base + inner);
            vv = *reinterpret_cast<const uint4*>(base + 2 * inner);
        }
        *reinterpret_cast<uint4*>(sK + r * SB_LD + c) = kv;
        *reinterpret_cast<uint4*>(sV + r * SB_LD + c) = vv;
    }

    float dk[4][4], dv[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dk[j][e] = 0.f; dv[j][e] = 0.f; }

    const int wm = warp & 3;      // 16-row slab of the query block (S / dP / dQ)
    const int wn = warp >> 2;     // 32-key quarter (S / dP), 16-dim quarter (dQ)
    const int wk = warp & 7;      // 16-key slab of dK / dV
    const int wd = warp >> 3;     // 32-dim half of dK / dV
    const int q_blocks = (tokens + SB_QB - 1) / SB_QB;

    for (int qb = 0; qb < q_blocks; ++qb) {
        const int q0 = qb * SB_QB;
        __syncthreads();   // previous iteration's readers of sQ / sdO / sP / sdS are done (also orders the K/V fill)
        // ---- Q / dO block -> smem, D = rowsum(dO o O), lse ----
        for (int idx = tid; idx < SB_QB * 8; idx += SB_THREADS) {
            const int r = idx >> 3, c = (idx & 7) * 8;
            uint4 qv = make_uint4(0u, 0u, 0u, 0u), dov = qv;
            if (q0 + r < tokens) {
                qv = *reinterpret_cast<const uint4*>(qkv + (row0 + q0 + r) * (3 * inner) + h * SB_DH + c);
                dov = *reinterpret_cast<const uint4*>(dout + (row0 + q0 + r) * inner + h * SB_DH + c);
            }
            *reinterpret_cast<uint4*>(sQ + r * SB_LD + c) = qv;
            *reinterpret_cast<uint4*>(sdO + r * SB_LD + c) = dov;
        }
        {
            const int r = tid >> 3, part = tid & 7;    // 8 threads per row, 8 dims each
            float acc = 0.f;
            if (q0 + r < tokens) {
                float a[8], b[8];
                load8(dout + (row0 + q0 + r) * inner + h * SB_DH + part * 8, a);
                load8(o + (row0 + q0 + r) * inner + h * SB_DH + part * 8, b);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc = fmaf(a[e], b[e], acc);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            if (part == 0) {
                sD[r] = acc;
                sL[r] = (q0 + r < tokens) ? lse[(static_cast<int64_t>(bf) * heads + h) * tokens + q0 + r] : INFINITY;
            }
        }
        __syncthreads();

        // ---- S = Q K^T and dP = dO V^T for this warp's 16 x 64 sub-tile ----
        float s[4][4], dp[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) { s[j][e] = 0.f; dp[j][e] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t aq[4], ao[4];
            ldsm_x4(aq, sQ + (wm * 16 + (lmat & 1) * 8 + lr) * SB_LD + kk * 16 + (lmat >> 1) * 8);
            ldsm_x4(ao, sdO + (wm * 16 + (lmat & 1) * 8 + lr) * SB_LD + kk * 16 + (lmat >> 1) * 8);
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {       // two 8-key n-tiles per ldmatrix.x4
                uint32_t bk[4], bv[4];
                const int krow = wn * 32 + jp * 16 + (lmat >> 1) * 8 + lr;
                ldsm_x4(bk, sK + krow * SB_LD + kk * 16 + (lmat & 1) * 8);
                ldsm_x4(bv, sV + krow * SB_LD + kk * 16 + (lmat & 1) * 8);
                mma16816(s[2 * jp], aq, bk[0], bk[1]);
                mma16816(s[2 * jp + 1], aq, bk[2], bk[3]);
                mma16816(dp[2 * jp], ao, bv[0], bv[1]);
                mma16816(dp[2 * jp + 1], ao, bv[2], bv[3]);
            }
        }
        // ---- P, dS -> smem (bf16) ----
        {
            const int r0 = wm * 16 + g, r1 = r0 + 8;
            const float l0 = sL[r0], l1 = sL[r1];
            const float d0 = sD[r0], d1 = sD[r1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kl = wn * 32 + j * 8 + 2 * t;          // key inside the block
                const bool ok0 = key0 + kl < tokens, ok1 = key0 + kl + 1 < tokens;
                float p00 = ok0 ? exp2f(fmaf(s[j][0], scale_log2, -l0)) : 0.f;
                float p01 = ok1 ? exp2f(fmaf(s[j][1], scale_log2, -l0)) : 0.f;
                float p10 = ok0 ? exp2f(fmaf(s[j][2], scale_log2, -l1)) : 0.f;
                float p11 = ok1 ? exp2f(fmaf(s[j][3], scale_log2, -l1)) : 0.f;
                *reinterpret_cast<uint32_t*>(sP + r0 * SB_LDP + kl) = pkbf(p00, p01);
                *reinterpret_cast<uint32_t*>(sP + r1 * SB_LDP + kl) = pkbf(p10, p11);
                *reinterpret_cast<uint32_t*>(sdS + r0 * SB_LDP + kl) =
                    pkbf(p00 * (dp[j][0] - d0) * scale, p01 * (dp[j][1] - d0) * scale);
                *reinterpret_cast<uint32_t*>(sdS + r1 * SB_LDP + kl) =
                    pkbf(p10 * (dp[j][2] - d1) * scale, p11 * (dp[j][3] - d1) * scale);
                if (cam != nullptr) {
                    // relevance pass: cam[frame, q, k] += relu(dA o A) / heads, dA = dP (gradient w.r.t. the probabilities)
                    const float ih = 1.0f / static_cast<float>(heads);
                    const int kg = key0 + kl;
                    float* c0 = cam + (static_cast<int64_t>(bf) * tokens + q0 + r0) * tokens + kg;
                    float* c1 = cam + (static_cast<int64_t>(bf) * tokens + q0 + r1) * tokens + kg;
                    if (q0 + r0 < tokens) {
                        if (ok0) atomicAdd(c0, fmaxf(p00 * dp[j][0], 0.f) * ih);
                        if (ok1) atomicAdd(c0 + 1, fmaxf(p01 * dp[j][1], 0.f) * ih);
                    }
                    if (q0 + r1 < tokens) {
                        if (ok0) atomicAdd(c1, fmaxf(p10 * dp[j][2], 0.f) * ih);
                        if (ok1) atomicAdd(c1 + 1, fmaxf(p11 * dp[j][3], 0.f) * ih);
                    }
                }
            }
        }
        __syncthreads();

        // ---- dV += P^T dO, dK += dS^T Q: this warp owns keys [16 wk, 16 wk + 16) x dims [32 wd, 32 wd + 32) ----
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {           // 16 queries per step
            uint32_t ap[4], as_[4];
            // A[m = key][k = query] = X[query][key]: transposed 8x8 loads
            ldsm_x4_t(ap, sP + (kk * 16 + (lmat >> 1) * 8 + lr) * SB_LDP + wk * 16 + (lmat & 1) * 8);
            ldsm_x4_t(as_, sdS + (kk * 16 + (lmat >> 1) * 8 + lr) * SB_LDP + wk * 16 + (lmat & 1) * 8);
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {       // two 8-dim n-tiles per ldmatrix.x4
                uint32_t bo[4], bq[4];
                // B[k = query][n = dim] = X[query][dim]: rows are k -> transposed loads
                ldsm_x4_t(bo, sdO + (kk * 16 + (lmat & 1) * 8 + lr) * SB_LD + wd * 32 + jp * 16 + (lmat >> 1) * 8);
                ldsm_x4_t(bq, sQ + (kk * 16 + (lmat & 1) * 8 + lr) * SB_LD + wd * 32 + jp * 16 + (lmat >> 1) * 8);
                mma16816(dv[2 * jp], ap, bo[0], bo[1]);
                mma16816(dv[2 * jp + 1], ap, bo[2], bo[3]);
                mma16816(dk[2 * jp], as_, bq[0], bq[1]);
                mma16816(dk[2 * jp + 1], as_, bq[2], bq[3]);
            }
        }
        // ---- dQ partial = dS K for rows [16 wm, +16) x dims [16 wn, +16), accumulated across key blocks ----
        {
            float dq[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) dq[j][e] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {       // 16 keys per step
                uint32_t a[4], b[4];
                ldsm_x4(a, sdS + (wm * 16 + (lmat & 1) * 8 + lr) * SB_LDP + kk * 16 + (lmat >> 1) * 8);
                ldsm_x4_t(b, sK + (kk * 16 + (lmat & 1) * 8 + lr) * SB_LD + wn * 16 + (lmat >> 1) * 8);
                mma16816(dq[0], a, b[0], b[1]);
                mma16816(dq[1], a, b[2], b[3]);
            }
            const int r0 = q0 + wm * 16 + g, r1 = r0 + 8;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int d = h * SB_DH + wn * 16 + j * 8 + 2 * t;
                if (r0 < tokens) {
                    float* dst = dq_acc + (row0 + r0) * inner + d;
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(dq[j][0]), "f"(dq[j][1]) : "memory");
                }
                if (r1 < tokens) {
                    float* dst = dq_acc + (row0 + r1) * inner + d;
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(dq[j][2]), "f"(dq[j][3]) : "memory");
                }
            }
        }
    }

    // ---- dK, dV -> dqkv (k columns at inner, v columns at 2 inner) ----
    {
        const int r0 = key0 + wk * 16 + g, r1 = r0 + 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = h * SB_DH + wd * 32 + j * 8 + 2 * t;
            if (r0 < tokens) {
                bf16* base = dqkv + (row0 + r0) * (3 * inner) + d;
                *reinterpret_cast<uint32_t*>(base + inner) = pkbf(dk[j][0], dk[j][1]);
                *reinterpret_cast<uint32_t*>(base + 2 * inner) = pkbf(dv[j][0], dv[j][1]);
            }
            if (r1 < tokens) {
                bf16* base = dqkv + (row0 + r1) * (3 * inner) + d;
                *reinterpret_cast<uint32_t*>(base + inner) = pkbf(dk[j][2], dk[j][3]);
                *reinterpret_cast<uint32_t*>(base + 2 * inner) = pkbf(dv[j][2], dv[j][3]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// tcgen05 version.  Legacy mma.sync peaks near 512 FLOP/clk/SM on sm_100a (the kernel above sits at that ceiling,
// 112-126 TFLOP/s), so the five contractions move to the 5th-gen tensor cores:
//   one CTA = (frame, head, 128-key chunk); loop over the 128-query tiles.
//   TMEM : S [128q x 128k] cols 0-127, dP cols 128-255, dV [128k x 64] 256-319, dK 320-383, dQ [128q x 64] 384-447.
//   smem : K, V chunk (16 KB each), Q / dO tile double buffered (2 x 32 KB), P and dS [128q x 128k] bf16 (32 KB each),
//          all 128-byte swizzled.  One byte layout serves every operand role: a [rows x 64] tile is K-major for
//          S = Q K^T / dP = dO V^T and MN-major (N = 64 contiguous) as the B of dV = P^T dO, dK = dS^T Q, dQ = dS K;
//          the P / dS tile [atom(64 keys)][query row][128 B] is K-major A for dQ and MN-major A (M = keys) for dV / dK.
//   roles: warp 0 TMA, warp 1 MMA issue, warp 2 TMEM alloc, warps 4-11: D = rowsum(dO o O), P = exp2(S c - lse),
//          dS = P o (dP - D) * scale -> smem, dQ tile -> red.global.add (summed over the key chunks), final dK / dV.
// ------------------------------------------------------------------------------------------
constexpr int TB_PARTS = 4;                           // softmax threads per query row (32 of the 128 key columns each)
constexpr int TB_SM_WARPS = 4 * TB_PARTS;             // 16 (8 until r6r: 64 columns per thread at 168 registers)
constexpr int TB_THREADS = 128 + 32 * TB_SM_WARPS;
constexpr int TB_TILE = 128 * 64 * 2;                 // 16 KB: 128 rows x 64 bf16
// K / V chunk double buffered by item parity, Q / dO tile double buffered by tile parity, P and dS single
constexpr int TB_K_OFF = 0, TB_V_OFF = 2 * TB_TILE, TB_Q_OFF = 4 * TB_TILE, TB_DO_OFF = 6 * TB_TILE;
constexpr int TB_P_OFF = 8 * TB_TILE, TB_DS_OFF = 10 * TB_TILE, TB_MISC_OFF = 12 * TB_TILE;      // 192 KB
constexpr int TB_SMEM = TB_MISC_OFF + 1024 + 256;

// PERSISTENT (r6s): one CTA per SM walks items = (frame, head, 128-key chunk); the producer fetches the next item's
// K / V chunk and first Q / dO tile while the current item computes, TMEM is allocated once, and the dK / dV epilogue
// of item n overlaps the S / dP MMAs of item n + 1.  The one-item-per-CTA version paid launch, barrier set-up, TMEM
// allocation and the first TMA round trip 10 752 times per layer: 24.8 us per item against ~12 us of work.
__device__ __forceinline__ int tb_item(int n, int k_chunks, int grouped) {
    if (!grouped) return static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
    const int grp = n / k_chunks;
    return (static_cast<int>(blockIdx.x) + grp * static_cast<int>(gridDim.x)) * k_chunks + (n - grp * k_chunks);
}

__global__ void __launch_bounds__(TB_THREADS, 1)
attn_spatial_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                           const float* __restrict__ d_rows, const float* __restrict__ lse,
                           bf16* __restrict__ dqkv, float* __restrict__ dq_acc, float* __restrict__ cam, int tokens,
                           int heads, int items, float scale, int grouped, int no_red) {
    extern __shared__ uint8_t tb_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tb_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TB_MISC_OFF);
    uint64_t* kv_full = bars;         // [2]
    uint64_t* kv_empty = bars + 2;    // [2]
    uint64_t* q_full = bars + 4;      // [2]
    uint64_t* q_empty = bars + 6;     // [2]
    uint64_t* s_full = bars + 8;
    uint64_t* p_ready = bars + 9;
    uint64_t* dq_full = bars + 10;
    uint64_t* dq_empty = bars + 11;
    uint64_t* fin = bars + 12;        // dK / dV of the item complete in TMEM
    uint64_t* dkv_empty = bars + 13;  // ... and read out by the epilogue warps
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_chunks = (tokens + 127) / 128;
    const int q_tiles = k_chunks;
    const int inner = heads * SB_DH;
    const float scale_log2 = scale * 1.4426950408889634f;
    // grouped: a CTA takes the k_chunks key chunks of one (frame, head) back to back, so the three read-modify-writes of
    // that (frame, head)'s dQ rows and the three reads of its Q / dO tiles hit in L2 instead of going to HBM
    const int groups = items / k_chunks;
    const int my_groups = groups > static_cast<int>(blockIdx.x)
                              ? (groups - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                              : 0;
    const int my_items = grouped ? my_groups * k_chunks
                                 : (items > static_cast<int>(blockIdx.x)
                                        ? (items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                                        : 0);

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_qkv); tma_prefetch_desc(&tm_do); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(kv_full + i, 1); mbar_init(kv_empty + i, 1);
            mbar_init(q_full + i, 1); mbar_init(q_empty + i, 1);
        }
        mbar_init(s_full, 1); mbar_init(p_ready, TB_SM_WARPS); mbar_init(dq_full, 1); mbar_init(dq_empty, TB_SM_WARPS);
        mbar_init(fin, 1); mbar_init(dkv_empty, TB_SM_WARPS);
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc(tmem_holder, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    const uint32_t t_s = tmem, t_dp = tmem + 128, t_dv = tmem + 256, t_dk = tmem + 320, t_dq = tmem + 384;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int n = 0; n < my_items; ++n) {
                const int item = tb_item(n, k_chunks, grouped);
                const int kc = item % k_chunks, h = (item / k_chunks) % heads, bf = item / (k_chunks * heads);
                const int ks = n & 1;
                mbar_wait_sleep(kv_empty + ks, ((n >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(kv_full + ks, 2 * TB_TILE);
                tma_load_3d(smem + TB_K_OFF + ks * TB_TILE, &tm_qkv, kv_full + ks, inner + h * SB_DH, kc * 128, bf);
                tma_load_3d(smem + TB_V_OFF + ks * TB_TILE, &tm_qkv, kv_full + ks, 2 * inner + h * SB_DH, kc * 128, bf);
                for (int t = 0; t < q_tiles; ++t) {
                    const int g = n * q_tiles + t;
                    const int slot = g & 1;
                    mbar_wait_sleep(q_empty + slot, ((g >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(q_full + slot, 2 * TB_TILE);
                    tma_load_3d(smem + TB_Q_OFF + slot * TB_TILE, &tm_qkv, q_full + slot, h * SB_DH, t * 128, bf);
                    tma_load_3d(smem + TB_DO_OFF + slot * TB_TILE, &tm_do, q_full + slot, h * SB_DH, t * 128, bf);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            const uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);     // S, dP: A, B K-major
            const uint32_t id_t = make_idesc_bf16(128, 64, 1, 1);      // dV, dK: A (P^T / dS^T) and B MN-major
            const uint32_t id_q = make_idesc_bf16(128, 64, 0, 1);      // dQ: A K-major, B MN-major
            const uint64_t d_kmaj = make_smem_desc(0, 0, 1024, SWZ_128B);
            const uint64_t d_mn_a = make_smem_desc(0, 128 * 128, 1024, SWZ_128B);   // atoms of 64 keys are 16 KB apart
            const uint64_t d_mn_b = make_smem_desc(0, 128 * 128, 1024, SWZ_128B);   // N = 64: a single atom
            auto fld = [&](int off) { return static_cast<uint64_t>((smem_u32(smem + off) & 0x3FFFFu) >> 4); };
            const uint64_t p_f = fld(TB_P_OFF), ds_f = fld(TB_DS_OFF);
            for (int n = 0; n < my_items; ++n) {
                const int ks = n & 1;
                const uint64_t k_f = fld(TB_K_OFF + ks * TB_TILE), v_f = fld(TB_V_OFF + ks * TB_TILE);
                mbar_wait(kv_full + ks, (n >> 1) & 1);
                for (int t = 0; t < q_tiles; ++t) {
                    const int g = n * q_tiles + t;
                    const int slot = g & 1;
                    const uint64_t q_f = fld(TB_Q_OFF + slot * TB_TILE), do_f = fld(TB_DO_OFF + slot * TB_TILE);
                    mbar_wait(q_full + slot, (g >> 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k)      // S = Q K^T
                        umma_f16_ss(t_s, d_kmaj | (q_f + 2 * k), d_kmaj | (k_f + 2 * k), id_s, k != 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 4; ++k)      // dP = dO V^T
                        umma_f16_ss(t_dp, d_kmaj | (do_f + 2 * k), d_kmaj | (v_f + 2 * k), id_s, k != 0 ? 1u : 0u);
                    umma_commit(s_full);
                    mbar_wait(p_ready, g & 1);                       // P, dS in smem; S / dP TMEM consumed
                    mbar_wait(dq_empty, (g & 1) ^ 1);                // previous dQ tile drained
                    if (t == 0) mbar_wait(dkv_empty, (n & 1) ^ 1);   // previous item's dK / dV read out
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < 8; ++j)      // dQ = dS K   (A K-major: key atom j>>2, 32 B per k-step; B = K MN-major)
                        umma_f16_ss(t_dq, d_kmaj | (ds_f + (j >> 2) * (128 * 128 >> 4) + (j & 3) * 2),
                                    d_mn_b | (k_f + j * (16 * 128 >> 4)), id_q, j != 0 ? 1u : 0u);
                    umma_commit(dq_full);
#pragma unroll
                    for (int j = 0; j < 8; ++j)      // dV += P^T dO   (k-step = 16 queries = 2048 B in both operands)
                        umma_f16_ss(t_dv, d_mn_a | (p_f + j * (16 * 128 >> 4)), d_mn_b | (do_f + j * (16 * 128 >> 4)), id_t,
                                    (t | j) != 0 ? 1u : 0u);
#pragma unroll
                    for (int j = 0; j < 8; ++j)      // dK += dS^T Q
                        umma_f16_ss(t_dk, d_mn_a | (ds_f + j * (16 * 128 >> 4)), d_mn_b | (q_f + j * (16 * 128 >> 4)), id_t,
                                    (t | j) != 0 ? 1u : 0u);
                    umma_commit(q_empty + slot);
                }
                umma_commit(kv_empty + ks);
                umma_commit(fin);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ================= softmax / dS / epilogues =================
        const int quad = warp & 3, part = (warp - 4) >> 2;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
        const float ih = 1.0f / static_cast<float>(heads);
        for (int n = 0; n < my_items; ++n) {
        const int item = tb_item(n, k_chunks, grouped);
        const int kc = item % k_chunks, h = (item / k_chunks) % heads, bf = item / (k_chunks * heads);
        const int64_t row0 = static_cast<int64_t>(bf) * tokens;
        for (int t = 0; t < q_tiles; ++t) {
            const int g = n * q_tiles + t;
            const int q_idx = t * 128 + row;
            const bool row_ok = q_idx < tokens;
            // D = rowsum(dO o O) of this query row and head: computed once per row by attn_spatial_bwd_d_kernel into the
            // (not yet written) dQ columns of dqkv.  The first version recomputed it in every key-chunk CTA from 2 x 64
            // bytes of uncoalesced global loads per thread and exchanged halves through shared memory + a named barrier —
            // 11 % of the stall samples (profiles/r3s_attn_spatial_bwd_ncu_source.txt).
            // (Loading these one tile ahead measured SLOWER — 1.476 vs 1.356 ms per launch, r6w: the two extra live values
            //  and the item arithmetic cost more than the L2 round trip they hide.)
            const float dsum = row_ok ? __ldg(d_rows + (row0 + q_idx) * (3 * inner / 2) + h) : 0.f;
            const float l = row_ok ? __ldg(lse + (static_cast<int64_t>(bf) * heads + h) * tokens + q_idx) : INFINITY;

            mbar_wait(s_full, g & 1);
            tc_fence_after();
            {
                const int col0 = part * 32;                        // key column inside the chunk
                uint32_t rs[32], rd[32];
                tmem_ld_32x32b_x32(t_s + lane_base + col0, rs);
                tmem_ld_32x32b_x32(t_dp + lane_base + col0, rd);
                tmem_ld_wait();
                uint32_t pk[16], dk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float p[2], ds[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int kg = kc * 128 + col0 + 2 * j + e;
                        const float dpv = __uint_as_float(rd[2 * j + e]);
                        p[e] = (kg < tokens) ? ex2_approx(fmaf(__uint_as_float(rs[2 * j + e]), scale_log2, -l)) : 0.f;
                        ds[e] = p[e] * (dpv - dsum) * scale;
                        if (cam != nullptr && row_ok && kg < tokens)
                            atomicAdd(cam + (static_cast<int64_t>(bf) * tokens + q_idx) * tokens + kg,
                                      fmaxf(p[e] * dpv, 0.f) * ih);
                    }
                    pk[j] = pack_bf16x2(p[0], p[1]);
                    dk[j] = pack_bf16x2(ds[0], ds[1]);
                }
                // [atom = 64 keys][query row][128 B], 16-byte chunks XOR-swizzled by (row & 7)
                const int atom = col0 >> 6, chunk0 = (col0 & 63) >> 3;
                uint8_t* prow = smem + TB_P_OFF + atom * (128 * 128) + row * 128;
                uint8_t* drow = smem + TB_DS_OFF + atom * (128 * 128) + row * 128;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int chunk = (chunk0 + g) ^ (row & 7);
                    *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                    *reinterpret_cast<uint4*>(drow + chunk * 16) = make_uint4(dk[4 * g], dk[4 * g + 1], dk[4 * g + 2], dk[4 * g + 3]);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);

            // dQ tile: this thread's 32 of the 64 dims of its query row, accumulated over the key chunks in global
            mbar_wait(dq_full, g & 1);
            tc_fence_after();
            {
                uint32_t r[16];
                tmem_ld_32x32b_x16(t_dq + lane_base + part * 16, r);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dq_empty);
                if (row_ok && !no_red) {
                    float* dst = dq_acc + (row0 + q_idx) * inner + h * SB_DH + part * 16;
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * g),
                                     "f"(__uint_as_float(r[4 * g])), "f"(__uint_as_float(r[4 * g + 1])),
                                     "f"(__uint_as_float(r[4 * g + 2])), "f"(__uint_as_float(r[4 * g + 3]))
                                     : "memory");
                }
            }
        }
        // ---- dK, dV of this key chunk ----
        mbar_wait(fin, n & 1);
        tc_fence_after();
        {
            const int key = kc * 128 + row;
            uint32_t rk[16], rv[16];
            tmem_ld_32x32b_x16(t_dk + lane_base + part * 16, rk);
            tmem_ld_32x32b_x16(t_dv + lane_base + part * 16, rv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dkv_empty);      // the next item's dV / dK MMAs may overwrite the accumulators
            if (key < tokens) {
                bf16* base = dqkv + (row0 + key) * (3 * inner) + h * SB_DH + part * 16;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint4 a, b;
                    a.x = pack_bf16x2(__uint_as_float(rk[8 * g]), __uint_as_float(rk[8 * g + 1]));
                    a.y = pack_bf16x2(__uint_as_float(rk[8 * g + 2]), __uint_as_float(rk[8 * g + 3]));
                    a.z = pack_bf16x2(__uint_as_float(rk[8 * g + 4]), __uint_as_float(rk[8 * g + 5]));
                    a.w = pack_bf16x2(__uint_as_float(rk[8 * g + 6]), __uint_as_float(rk[8 * g + 7]));
                    b.x = pack_bf16x2(__uint_as_float(rv[8 * g]), __uint_as_float(rv[8 * g + 1]));
                    b.y = pack_bf16x2(__uint_as_float(rv[8 * g + 2]), __uint_as_float(rv[8 * g + 3]));
                    b.z = pack_bf16x2(__uint_as_float(rv[8 * g + 4]), __uint_as_float(rv[8 * g + 5]));
                    b.w = pack_bf16x2(__uint_as_float(rv[8 * g + 6]), __uint_as_float(rv[8 * g + 7]));
                    *reinterpret_cast<uint4*>(base + inner + 8 * g) = a;
                    *reinterpret_cast<uint4*>(base + 2 * inner + 8 * g) = b;
                }
            }
        }
        }   // items
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// D[row, head] = sum_d dO[row, head, d] * O[row, head, d] (fp32), one warp per token row, written to the first `heads`
// floats of the row's dQ columns in dqkv — scratch until attn_spatial_bwd_dq_kernel overwrites them at the end.
__global__ void __launch_bounds__(256)
attn_spatial_bwd_d_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, bf16* __restrict__ dqkv, int64_t rows,
                          int heads) {
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int inner = heads * SB_DH;
    float* dst = reinterpret_cast<float*>(dqkv + row * (3 * inner));
    for (int c0 = 0; c0 < inner; c0 += 256) {          // 32 lanes x 8 elements: 4 heads per pass, 8 lanes per head
        const int c = c0 + lane * 8;
        float acc = 0.f;
        if (c < inner) {
            float a[8], b[8];
            load8(dout + row * inner + c, a);
            load8(o + row * inner + c, b);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc = fmaf(a[e], b[e], acc);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if ((lane & 7) == 0 && c < inner) dst[c >> 6] = acc;
    }
}

// dq_acc fp32 [rows, inner] -> q columns of dqkv [rows, 3 inner] (bf16)
__global__ void __launch_bounds__(256)
attn_spatial_bwd_dq_kernel(const float* __restrict__ dq_acc, bf16* __restrict__ dqkv, int64_t rows, int inner) {
    const int c8 = inner >> 3;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= rows * c8) return;
    const int64_t r = idx / c8;
    const int c = static_cast<int>(idx - r * c8) * 8;
    float v[8];
    load8(dq_acc + r * inner + c, v);
    store8(dqkv + r * (3 * inner) + c, v);
}

}  // namespace istvt

using namespace istvt;

// qkv: bf16 [batch_frames*tokens, 3*heads*64]; o, dout: bf16 [rows, heads*64]; lse: fp32 [batch_frames, heads, tokens]
// (log2 domain, from istvt_attn_spatial_fwd_lse); dqkv: bf16 [rows, 3*heads*64] (fully written);
// dq_scratch: fp32 [rows, heads*64] workspace (zero-filled by this call).
static int attn_spatial_bwd_launch(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv,
                                   float* dq_scratch, float* cam, int batch_frames, int tokens, int heads, float scale,
                                   istvt_stream_t stream) {
    ISTVT_REQUIRE(qkv && o && dout && lse && dqkv && dq_scratch);
    ISTVT_REQUIRE(batch_frames > 0 && tokens > 0 && heads > 0);
    ISTVT_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(dout) |
                    reinterpret_cast<uintptr_t>(dqkv) | reinterpret_cast<uintptr_t>(dq_scratch)) & 15) == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int inner = heads * SB_DH;
    const int64_t rows = static_cast<int64_t>(batch_frames) * tokens;
    ISTVT_CHECK_CUDA(cudaMemsetAsync(dq_scratch, 0, static_cast<size_t>(rows) * inner * sizeof(float), st));
    static const bool legacy = []() { const char* e = getenv("ISTVT_ATTN_BWD_LEGACY"); return e && atoi(e) != 0; }();
    const int k_blocks = (tokens + SB_KB - 1) / SB_KB;
    const int64_t grid = static_cast<int64_t>(batch_frames) * heads * k_blocks;
    ISTVT_REQUIRE(grid < (int64_t(1) << 31));
    if (legacy || tokens > 384) {
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM));
        attn_spatial_bwd_kernel<<<static_cast<unsigned>(grid), SB_THREADS, SB_SMEM, st>>>(
            static_cast<const bf16*>(qkv), static_cast<const bf16*>(o), static_cast<const bf16*>(dout), lse,
            static_cast<bf16*>(dqkv), dq_scratch, cam, tokens, heads, scale);
    } else {
        CUtensorMap tm_qkv, tm_do;
        {
            const uint64_t dims[3] = {static_cast<uint64_t>(3 * inner), static_cast<uint64_t>(tokens),
                                      static_cast<uint64_t>(batch_frames)};
            const uint64_t strides[2] = {static_cast<uint64_t>(3 * inner) * 2, static_cast<uint64_t>(tokens) * 3 * inner * 2};
            const uint32_t box[3] = {SB_DH, 128, 1};
            int rc = encode_tmap(&tm_qkv, qkv, ISTVT_BF16, 3, dims, strides, box, 3);
            if (rc != ISTVT_OK) return rc;
        }
        {
            const uint64_t dims[3] = {static_cast<uint64_t>(inner), static_cast<uint64_t>(tokens),
                                      static_cast<uint64_t>(batch_frames)};
            const uint64_t strides[2] = {static_cast<uint64_t>(inner) * 2, static_cast<uint64_t>(tokens) * inner * 2};
            const uint32_t box[3] = {SB_DH, 128, 1};
            int rc = encode_tmap(&tm_do, dout, ISTVT_BF16, 3, dims, strides, box, 3);
            if (rc != ISTVT_OK) return rc;
        }
        attn_spatial_bwd_d_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(
            static_cast<const bf16*>(o), static_cast<const bf16*>(dout), static_cast<bf16*>(dqkv), rows, heads);
        count_launch();
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM));
        // ISTVT_SAB_GROUPED=0: items round-robin over the CTAs instead of one (frame, head) per CTA at a time (A/B);
        // ISTVT_SAB_NORED=1 (timing experiment, wrong dQ): without the red.global accumulation of dQ
        static const int grouped_env = []() { const char* e = getenv("ISTVT_SAB_GROUPED"); return e ? atoi(e) : 1; }();
        static const int nored_env = []() { const char* e = getenv("ISTVT_SAB_NORED"); return e ? atoi(e) : 0; }();
        const int64_t ctas = grid < sm_count() ? grid : sm_count();       // persistent: one CTA per SM walks the items
        attn_spatial_bwd_tc_kernel<<<static_cast<unsigned>(ctas), TB_THREADS, TB_SMEM, st>>>(
            tm_qkv, tm_do, reinterpret_cast<const float*>(dqkv), lse, static_cast<bf16*>(dqkv), dq_scratch, cam, tokens,
            heads, static_cast<int>(grid), scale, grouped_env, nored_env);
    }
    count_launch();
    const int64_t n = rows * (inner / 8);
    attn_spatial_bwd_dq_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(dq_scratch,
                                                                                       static_cast<bf16*>(dqkv), rows, inner);
    count_launch();
    return launch_status();
}

extern "C" int istvt_attn_spatial_bwd(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv,
                                      float* dq_scratch, int batch_frames, int tokens, int heads, float scale,
                                      istvt_stream_t stream) {
    return attn_spatial_bwd_launch(qkv, o, dout, lse, dqkv, dq_scratch, nullptr, batch_frames, tokens, heads, scale, stream);
}

// Same backward, additionally accumulating the head-averaged gradient-weighted attention of the relevance pass:
// cam[batch_frames, tokens, tokens] (fp32, zero-filled by the caller) += relu(dA o A) / heads.
extern "C" int istvt_attn_spatial_bwd_cam(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv,
                                          float* dq_scratch, float* cam, int batch_frames, int tokens, int heads,
                                          float scale, istvt_stream_t stream) {
    ISTVT_REQUIRE(cam != nullptr);
    return attn_spatial_bwd_launch(qkv, o, dout, lse, dqkv, dq_scratch, cam, batch_frames, tokens, heads, scale, stream);
}
