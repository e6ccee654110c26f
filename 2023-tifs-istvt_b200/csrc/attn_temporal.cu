// Temporal self-attention across frames at each token position (network/vivit/module.py:197-205).
//
// Problem shape per (clip b, head h, position p): F x F scores with head dim 64, F = T+1 = 7 (33 for the
// long-clip config) — far too small for the tensor cores, and HBM-bound: the kernel's job is to read
// q/k/v exactly once, IN PLACE (the reference makes three permute copies), and write the output rows.
// Frames of one position are `tokens` rows apart in the [rows, heads*64] projection buffers; a row
// holds all heads contiguously, so one CTA = one (b, p) with all heads: every row it touches is a
// full 1 KB (bf16) contiguous read.  K and V for the CTA are staged in shared memory, each thread
// owns one (head, query frame) pair with q and the output accumulator in registers and runs an
// online softmax over the F key frames (all lanes of a warp with the same head read K/V by broadcast;
// the +8 element row padding keeps different heads on different banks).
#include "common.cuh"
#include "ptx.cuh"
#include "simt_util.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

namespace istvt {

constexpr int TA_DH = 64;       // head dim
constexpr int TA_PAD = 8;       // smem row padding (elements)
constexpr int TA_ROW = TA_DH + TA_PAD;

template <typename T>
__global__ void __launch_bounds__(288)
attn_temporal_kernel(const T* __restrict__ qk, const T* __restrict__ v, T* __restrict__ out,
                     float* __restrict__ probs, int frames, int tokens, int heads, float scale) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    T* sk = reinterpret_cast<T*>(smem_raw);                 // [frames][heads][TA_ROW]
    T* sv = sk + static_cast<size_t>(frames) * heads * TA_ROW;

    const int b = blockIdx.x / tokens;
    const int pos = blockIdx.x - b * tokens;
    const int inner = heads * TA_DH;                          // 512
    const int64_t row0 = (static_cast<int64_t>(b) * frames) * tokens + pos;  // row of frame 0

    // ---- stage K and V: 8-element (16 B bf16 / 32 B fp32) chunks, coalesced along the row ----
    const int chunks_per_row = inner / 8;
    const int total_chunks = frames * chunks_per_row;
    for (int c = threadIdx.x; c < total_chunks; c += blockDim.x) {
        const int j = c / chunks_per_row;
        const int e = (c - j * chunks_per_row) * 8;  // element within the 512-wide row
        const int h = e / TA_DH;
        const int d = e - h * TA_DH;
        const int64_t row = row0 + static_cast<int64_t>(j) * tokens;
        float kv[8], vv[8];
        load8(qk + row * (2 * inner) + inner + e, kv);
        load8(v + row * inner + e, vv);
        store8(sk + (static_cast<size_t>(j) * heads + h) * TA_ROW + d, kv);
        store8(sv + (static_cast<size_t>(j) * heads + h) * TA_ROW + d, vv);
    }
    __syncthreads();

    const int t = threadIdx.x;
    if (t >= heads * frames) return;
    const int h = t / frames;
    const int i = t - h * frames;  // query frame
    const int64_t qrow = row0 + static_cast<int64_t>(i) * tokens;

    float q[TA_DH];
#pragma unroll
    for (int d = 0; d < TA_DH; d += 8) {
        float tmp[8];
        load8(qk + qrow * (2 * inner) + h * TA_DH + d, tmp);
#pragma unroll
        for (int e = 0; e < 8; ++e) q[d + e] = tmp[e] * scale;
    }

    float o[TA_DH];
#pragma unroll
    for (int d = 0; d < TA_DH; ++d) o[d] = 0.0f;
    float mx = -INFINITY, l = 0.0f;
    for (int j = 0; j < frames; ++j) {
        const T* kr = sk + (static_cast<size_t>(j) * heads + h) * TA_ROW;
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d < TA_DH; d += 8) {
            float kk[8];
            load8(kr + d, kk);
#pragma unroll
            for (int e = 0; e < 8; ++e) s = fmaf(q[d + e], kk[e], s);
        }
        const float mnew = fmaxf(mx, s);
        const float corr = expf(mx - mnew);  // 0 on the first key (mx = -inf)
        const float pj = expf(s - mnew);
        l = l * corr + pj;
        const T* vr = sv + (static_cast<size_t>(j) * heads + h) * TA_ROW;
#pragma unroll
        for (int d = 0; d < TA_DH; d += 8) {
            float vv[8];
            load8(vr + d, vv);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[d + e] = fmaf(o[d + e], corr, pj * vv[e]);
        }
        mx = mnew;
    }
    const float inv = 1.0f / l;
#pragma unroll
    for (int d = 0; d < TA_DH; d += 8) {
        float tmp[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) tmp[e] = o[d + e] * inv;
        store8(out + qrow * inner + h * TA_DH + d, tmp);
    }

    if (probs != nullptr) {
        // probs[b, h, pos, i, j]
        float* pr = probs + ((((static_cast<int64_t>(b) * heads + h) * tokens + pos) * frames) + i) * frames;
        for (int j = 0; j < frames; ++j) {
            const T* kr = sk + (static_cast<size_t>(j) * heads + h) * TA_ROW;
            float s = 0.0f;
#pragma unroll
            for (int d = 0; d < TA_DH; d += 8) {
                float kk[8];
                load8(kr + d, kk);
#pragma unroll
                for (int e = 0; e < 8; ++e) s = fmaf(q[d + e], kk[e], s);
            }
            pr[j] = expf(s - mx) * inv;
        }
    }
}

// ------------------------------------------------------------------------------------------
// bf16 production kernel for F <= 8 frames (the ISTVT configuration, F = 7): one WARP per
// (clip, position, head) on the legacy warp-level tensor-core path (mma.sync m16n8k16) — the
// problem (7x7x64) is far below a tcgen05 tile, but mma.sync consumes the bf16 rows exactly as
// they sit in HBM, so the kernel is ~100 instructions per warp with every global access a 16-byte
// vector: 2+2 for Q/K row g (lane = 4g+t), 2 for V rows 2t/2t+1, 2 stores.  Index tricks:
//   * the k (head-dim) order of Q.K^T is permuted identically for A and B, so each lane's two
//     16-byte loads feed the fragments directly (no shuffles);
//   * the n (head-dim) order of P.V is permuted (n = g  <->  dim 8g + nt) so that the B fragment
//     {V[2t][d], V[2t+1][d]} is a byte-permute of two 16-byte row loads and each lane ends up
//     owning 16 contiguous output dims of row g.
// A CTA (8 warps) covers all heads of one (clip, position): full 1 KB rows per request.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

__global__ void __launch_bounds__(256)
attn_temporal_mma_kernel(const __nv_bfloat16* __restrict__ qk, const __nv_bfloat16* __restrict__ v,
                         __nv_bfloat16* __restrict__ out, float* __restrict__ probs, int frames, int tokens,
                         int heads, float scale_log2, int64_t units) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2;          // fragment row: query frame (A / D), key frame (B of Q.K^T)
    const int t = lane & 3;
    const int64_t unit = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (unit >= units) return;
    const int h = static_cast<int>(unit % heads);
    const int64_t bp = unit / heads;
    const int64_t b = bp / tokens;
    const int pos = static_cast<int>(bp - b * tokens);
    const int inner = heads * TA_DH;
    const int64_t row0 = b * frames * tokens + pos;

    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    uint4 q0 = zero, q1 = zero, k0 = zero, k1 = zero, v0 = zero, v1 = zero;
    if (g < frames) {
        const __nv_bfloat16* qp = qk + (row0 + static_cast<int64_t>(g) * tokens) * (2 * inner) + h * TA_DH;
        q0 = ldg_nc_u4(qp + 8 * t);
        q1 = ldg_nc_u4(qp + 8 * (t + 4));
        k0 = ldg_nc_u4(qp + inner + 8 * t);
        k1 = ldg_nc_u4(qp + inner + 8 * (t + 4));
    }
    if (2 * t < frames)
        v0 = ldg_nc_u4(v + (row0 + static_cast<int64_t>(2 * t) * tokens) * inner + h * TA_DH + 8 * g);
    if (2 * t + 1 < frames)
        v1 = ldg_nc_u4(v + (row0 + static_cast<int64_t>(2 * t + 1) * tokens) * inner + h * TA_DH + 8 * g);

    // S[g][2t], S[g][2t+1] (rows 8-15 of the tile are padding)
    float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    mma_bf16_16816(s, q0.x, 0u, q0.y, 0u, k0.x, k0.y);
    mma_bf16_16816(s, q0.z, 0u, q0.w, 0u, k0.z, k0.w);
    mma_bf16_16816(s, q1.x, 0u, q1.y, 0u, k1.x, k1.y);
    mma_bf16_16816(s, q1.z, 0u, q1.w, 0u, k1.z, k1.w);

    const float s0 = (2 * t < frames) ? s[0] * scale_log2 : -INFINITY;
    const float s1 = (2 * t + 1 < frames) ? s[1] * scale_log2 : -INFINITY;
    float mx = fmaxf(s0, s1);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float p0, p1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(s0 - mx));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(s1 - mx));
    // the P.V MMA consumes bf16 P: normalise by the sum of the rounded values
    const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
    const uint32_t pp = *reinterpret_cast<const uint32_t*>(&pb);
    const float2 pf = __bfloat1622float2(pb);
    float sum = pf.x + pf.y;
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;

    if (probs != nullptr) {  // probs[b, h, pos, i, j], fp32, from the unrounded exponentials
        float sum32 = p0 + p1;
        sum32 += __shfl_xor_sync(0xffffffffu, sum32, 1);
        sum32 += __shfl_xor_sync(0xffffffffu, sum32, 2);
        const float inv32 = 1.0f / sum32;
        if (g < frames) {
            float* pr = probs + ((((b * heads + h) * tokens + pos) * frames) + g) * frames;
            if (2 * t < frames) pr[2 * t] = p0 * inv32;
            if (2 * t + 1 < frames) pr[2 * t + 1] = p1 * inv32;
        }
    }

    // O[g][16t + nt] and O[g][16t + 8 + nt] for nt = 0..7
    const uint32_t vw0[4] = {v0.x, v0.y, v0.z, v0.w};
    const uint32_t vw1[4] = {v1.x, v1.y, v1.z, v1.w};
    float olo[8], ohi[8];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const uint32_t b0 = __byte_perm(vw0[nt >> 1], vw1[nt >> 1], (nt & 1) ? 0x7632 : 0x5410);
        float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        mma_bf16_16816(d, pp, 0u, 0u, 0u, b0, 0u);
        olo[nt] = d[0] * inv;
        ohi[nt] = d[1] * inv;
    }
    if (g < frames) {
        __nv_bfloat16* op = out + (row0 + static_cast<int64_t>(g) * tokens) * inner + h * TA_DH + 16 * t;
        uint4 a, c;
        auto pk = [](float x, float y) {
            const __nv_bfloat162 r = __floats2bfloat162_rn(x, y);
            return *reinterpret_cast<const uint32_t*>(&r);
        };
        a.x = pk(olo[0], olo[1]); a.y = pk(olo[2], olo[3]); a.z = pk(olo[4], olo[5]); a.w = pk(olo[6], olo[7]);
        c.x = pk(ohi[0], ohi[1]); c.y = pk(ohi[2], ohi[3]); c.z = pk(ohi[4], ohi[5]); c.w = pk(ohi[6], ohi[7]);
        *reinterpret_cast<uint4*>(op) = a;
        *reinterpret_cast<uint4*>(op + 8) = c;
    }
}

// ------------------------------------------------------------------------------------------
// bf16 production kernels for 8 < F <= 48 frames (the long-clip configuration C5 has F = 33), the same
// one-warp-per-(clip, position, head), registers-only mma.sync scheme as above with the frame axis tiled:
//   * K is held as B fragments of Q.K^T for every 8-key n-tile (row 8 nj + g, 16-byte chunks t and t + 4), V as B
//     fragments of P.V for every 16-key k-step (rows 16 kk + 2t, + 1, + 8, + 9, chunk g; byte-permuted per n-tile);
//   * the query frames are walked in m16 tiles: Q rows 16 mi + g and + 8 (A fragments), S = 16 x 8 NT scores, softmax
//     over the n-tiles and the quad, and the score accumulators ARE the A fragments of P.V (c0,c1 / c2,c3 of n-tiles
//     2 kk and 2 kk + 1 = a0 / a1 and a2 / a3 of k-step kk) — no shuffles, no shared memory;
//   * P.V runs in fp16: with 33 keys the 8-bit mantissa of a bf16 P cost 2.9e-2 on the T = 32 golden logit (the SIMT
//     kernel it replaces: 1.4e-2); P in [0, 1] rounds to fp16 with 11 bits and V (bf16, 8 bits) converts exactly
//     (saturating at +-65504: V is a projection of LayerNorm output);
//   * legacy mma.sync peaks near 512 FLOP/clk/SM on sm_100a (profiles/README.md r1x), i.e. one m16n8k16 per 8 clk and
//     SM: the generic kernel's 132 MMAs per unit at F = 33 (3 m16 tiles x 5 n8 tiles, of which the third tile and the
//     fifth n-tile hold ONE frame) are 1056 clk per unit, above the unit's HBM time.  F = 16 MT + 1 (T = 16, 32: the
//     extra frame is the temporal class token) therefore has its own kernel: the 16 MT x 16 MT core on 64 MMAs
//     (MT = 2), the last KEY folded into the softmax / output of every row with fp32 FMAs on the lane's own fragments
//     (16 head dims per lane, quad reduction), and the last QUERY row done entirely on the FMA pipe from the K / V
//     fragments the lane already holds.
// HBM traffic = q, k, v once + the output once (4 KB per token row); 4 warps per CTA so that 3 CTAs overlap their
// load and compute phases on one SM.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                              uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float bf16lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf16hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t bf16x2_to_f16x2(uint32_t x) { return pack_f16x2(bf16lo(x), bf16hi(x)); }
__device__ __forceinline__ float2 unpack_f16x2(uint32_t x) {
    return __half22float2(*reinterpret_cast<const __half2*>(&x));
}
// partial dot product over the 16 head dims a lane holds (two 16-byte chunks of a bf16 row each)
__device__ __forceinline__ float dot16_bf16(const uint4 (&a)[2], const uint4 (&b)[2]) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const uint32_t aw[4] = {a[c].x, a[c].y, a[c].z, a[c].w}, bw[4] = {b[c].x, b[c].y, b[c].z, b[c].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc = fmaf(bf16lo(aw[i]), bf16lo(bw[i]), acc);
            acc = fmaf(bf16hi(aw[i]), bf16hi(bw[i]), acc);
        }
    }
    return acc;
}
__device__ __forceinline__ float quad_sum(float x) {
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    return x + __shfl_xor_sync(0xffffffffu, x, 2);
}
__device__ __forceinline__ float quad_max(float x) {
    x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
    return fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 2));
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float x, float y) {
    const __nv_bfloat162 r = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<const uint32_t*>(&r);
}

// TAIL = false: any 8 < F <= 16 MT (rows / keys >= F masked).  TAIL = true: F == 16 MT + 1 exactly.
// (4 CTAs per SM at 128 registers, ~0.5 KB of spills per thread, measured slower: 1.44 vs 1.20 ms per C5 step, r6g.)
template <int MT, bool TAIL>
__global__ void __launch_bounds__(128, TAIL ? 3 : 2)
attn_temporal_mma_wide_kernel(const __nv_bfloat16* __restrict__ qk, const __nv_bfloat16* __restrict__ v,
                              __nv_bfloat16* __restrict__ out, float* __restrict__ probs, int frames, int tokens,
                              int heads, float scale_log2, int64_t units) {
    constexpr int NT = 2 * MT;        // 8-key n-tiles of Q.K^T
    constexpr int R = 16 * MT;        // TAIL: index of the last frame
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2;
    const int t = lane & 3;
    const int64_t unit = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
    if (unit >= units) return;
    const int h = static_cast<int>(unit % heads);
    const int64_t bp = unit / heads;
    const int64_t b = bp / tokens;
    const int pos = static_cast<int>(bp - b * tokens);
    const int inner = heads * TA_DH;
    const int64_t row0 = b * frames * tokens + pos;
    const int64_t qk_pitch = static_cast<int64_t>(tokens) * (2 * inner);      // elements between consecutive frames
    const int64_t v_pitch = static_cast<int64_t>(tokens) * inner;
    const __nv_bfloat16* qbase = qk + row0 * (2 * inner) + h * TA_DH;
    const __nv_bfloat16* vbase = v + row0 * inner + h * TA_DH;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const int fmain = TAIL ? R : frames;       // frames covered by the MMA tiles

    // ---- K fragments: key 8 nj + g, head-dim chunks t and t + 4 ----
    uint4 kf[NT][2];
#pragma unroll
    for (int nj = 0; nj < NT; ++nj) {
        const int key = 8 * nj + g;
        kf[nj][0] = kf[nj][1] = zero;
        if (key < fmain) {
            const __nv_bfloat16* kp = qbase + key * qk_pitch + inner;
            kf[nj][0] = ldg_nc_u4(kp + 8 * t);
            kf[nj][1] = ldg_nc_u4(kp + 8 * (t + 4));
        }
    }
    // ---- V fragments: keys 16 kk + {2t, 2t+1, 2t+8, 2t+9}, head-dim chunk g ----
    uint4 vraw[MT][4];
#pragma unroll
    for (int kk = 0; kk < MT; ++kk)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int key = 16 * kk + 2 * t + (r & 1) + 8 * (r >> 1);
            vraw[kk][r] = zero;
            if (key < fmain) vraw[kk][r] = ldg_nc_u4(vbase + key * v_pitch + 8 * g);
        }
    // ---- TAIL: the last frame's q / k (this lane's 16 head dims) and v (this lane's 16 OUTPUT dims; chunk g) ----
    uint4 qt[2] = {zero, zero}, kt[2] = {zero, zero}, vt[2] = {zero, zero}, vtg = zero;
    if constexpr (TAIL) {
        const __nv_bfloat16* qp = qbase + R * qk_pitch;
        qt[0] = ldg_nc_u4(qp + 8 * t);         qt[1] = ldg_nc_u4(qp + 8 * (t + 4));
        kt[0] = ldg_nc_u4(qp + inner + 8 * t); kt[1] = ldg_nc_u4(qp + inner + 8 * (t + 4));
        vt[0] = ldg_nc_u4(vbase + R * v_pitch + 16 * t);
        vt[1] = ldg_nc_u4(vbase + R * v_pitch + 16 * t + 8);
        vtg = ldg_nc_u4(vbase + R * v_pitch + 8 * g);
    }
    // first query tile
    uint4 qa[2][2], qn[2][2];
    auto load_q = [&](int mi, uint4 (&q)[2][2]) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = 16 * mi + g + 8 * rr;
            q[rr][0] = q[rr][1] = zero;
            if (r < fmain) {
                q[rr][0] = ldg_nc_u4(qbase + r * qk_pitch + 8 * t);
                q[rr][1] = ldg_nc_u4(qbase + r * qk_pitch + 8 * (t + 4));
            }
        }
    };
    load_q(0, qa);

    // ---- B fragments of P.V in fp16: vb[kk][nt] = {V[16kk+2t][8g+nt], V[16kk+2t+1][.]}, {V[16kk+2t+8][.], V[16kk+2t+9][.]} ----
    uint32_t vb[MT][8][2];
#pragma unroll
    for (int kk = 0; kk < MT; ++kk) {
        const uint32_t w0[4] = {vraw[kk][0].x, vraw[kk][0].y, vraw[kk][0].z, vraw[kk][0].w};
        const uint32_t w1[4] = {vraw[kk][1].x, vraw[kk][1].y, vraw[kk][1].z, vraw[kk][1].w};
        const uint32_t w2[4] = {vraw[kk][2].x, vraw[kk][2].y, vraw[kk][2].z, vraw[kk][2].w};
        const uint32_t w3[4] = {vraw[kk][3].x, vraw[kk][3].y, vraw[kk][3].z, vraw[kk][3].w};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const uint32_t sel = (nt & 1) ? 0x7632 : 0x5410;
            vb[kk][nt][0] = bf16x2_to_f16x2(__byte_perm(w0[nt >> 1], w1[nt >> 1], sel));
            vb[kk][nt][1] = bf16x2_to_f16x2(__byte_perm(w2[nt >> 1], w3[nt >> 1], sel));
        }
    }
    float* pr_base = probs != nullptr ? probs + (((b * heads + h) * tokens + pos) * frames) * frames : nullptr;

    // ================= TAIL: the last query row, entirely on the FMA pipe =================
    if constexpr (TAIL) {
        float st[NT];                                 // S[R][8 nj + g] (log2 domain), the same in the 4 lanes of a quad
        float mx = -INFINITY;
#pragma unroll
        for (int nj = 0; nj < NT; ++nj) {
            st[nj] = quad_sum(dot16_bf16(qt, kf[nj])) * scale_log2;
            mx = fmaxf(mx, st[nj]);
        }
        const float stt = quad_sum(dot16_bf16(qt, kt)) * scale_log2;       // S[R][R]
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        mx = fmaxf(mx, stt);
        float sum = 0.f;
#pragma unroll
        for (int nj = 0; nj < NT; ++nj) {
            st[nj] = ex2f(st[nj] - mx);
            sum += st[nj];
        }
        const float ptt = ex2f(stt - mx);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        sum += ptt;
        const float inv = 1.0f / sum;
        if (pr_base != nullptr && t == 0) {
            float* pr = pr_base + static_cast<int64_t>(R) * frames;
#pragma unroll
            for (int nj = 0; nj < NT; ++nj) pr[8 * nj + g] = st[nj] * inv;
            if (g == 0) pr[R] = ptt * inv;
        }
        // O[R][8g + nt]: this lane's keys 16 kk + 2t + {0, 1, 8, 9}, then the quad (t) reduction and the last key
        float acc[8];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[nt] = 0.f;
#pragma unroll
        for (int kk = 0; kk < MT; ++kk) {
            // P[R][key] sits in every lane of quad g' = key & 7 as st[key >> 3]
            const float p0 = __shfl_sync(0xffffffffu, st[2 * kk], ((2 * t) << 2) | t);
            const float p1 = __shfl_sync(0xffffffffu, st[2 * kk], ((2 * t + 1) << 2) | t);
            const float p8 = __shfl_sync(0xffffffffu, st[2 * kk + 1], ((2 * t) << 2) | t);
            const float p9 = __shfl_sync(0xffffffffu, st[2 * kk + 1], ((2 * t + 1) << 2) | t);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float2 a = unpack_f16x2(vb[kk][nt][0]), c = unpack_f16x2(vb[kk][nt][1]);
                acc[nt] = fmaf(p0, a.x, fmaf(p1, a.y, fmaf(p8, c.x, fmaf(p9, c.y, acc[nt]))));
            }
        }
        const uint32_t vw[4] = {vtg.x, vtg.y, vtg.z, vtg.w};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            acc[nt] = quad_sum(acc[nt]);
            acc[nt] = fmaf(ptt, (nt & 1) ? bf16hi(vw[nt >> 1]) : bf16lo(vw[nt >> 1]), acc[nt]) * inv;
        }
        if (t == 0) {
            uint4 a;
            a.x = pack_bf16x2_rn(acc[0], acc[1]); a.y = pack_bf16x2_rn(acc[2], acc[3]);
            a.z = pack_bf16x2_rn(acc[4], acc[5]); a.w = pack_bf16x2_rn(acc[6], acc[7]);
            *reinterpret_cast<uint4*>(out + (row0 + static_cast<int64_t>(R) * tokens) * inner + h * TA_DH + 8 * g) = a;
        }
    }
    // the lane's 16 output dims of the last frame's V, for the last key's contribution to every row
    float vtf[16];
    if constexpr (TAIL) {
        const uint32_t w[8] = {vt[0].x, vt[0].y, vt[0].z, vt[0].w, vt[1].x, vt[1].y, vt[1].z, vt[1].w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            vtf[2 * i] = bf16lo(w[i]);
            vtf[2 * i + 1] = bf16hi(w[i]);
        }
    }

    // ================= the query tiles =================
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
        if (16 * mi >= fmain) break;                     // warp-uniform
        if (mi + 1 < MT && 16 * (mi + 1) < fmain) load_q(mi + 1, qn);      // prefetch the next tile's rows
        const int r0 = 16 * mi + g, r1 = r0 + 8;         // the two query frames of this lane
        // ---- S = Q K^T: s[nj] = {S[r0][8nj+2t], S[r0][8nj+2t+1], S[r1][8nj+2t], S[r1][8nj+2t+1]} ----
        float s[NT][4];
#pragma unroll
        for (int nj = 0; nj < NT; ++nj) {
            s[nj][0] = s[nj][1] = s[nj][2] = s[nj][3] = 0.f;
            if (8 * nj < fmain) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    mma_bf16_16816(s[nj], qa[0][c].x, qa[1][c].x, qa[0][c].y, qa[1][c].y, kf[nj][c].x, kf[nj][c].y);
                    mma_bf16_16816(s[nj], qa[0][c].z, qa[1][c].z, qa[0][c].w, qa[1][c].w, kf[nj][c].z, kf[nj][c].w);
                }
            }
        }
        float sl0 = -INFINITY, sl1 = -INFINITY;          // TAIL: scores against the last key (log2 domain)
        if constexpr (TAIL) {
            sl0 = quad_sum(dot16_bf16(qa[0], kt)) * scale_log2;
            sl1 = quad_sum(dot16_bf16(qa[1], kt)) * scale_log2;
        }
        // ---- softmax over the keys (n-tiles x quad lanes [+ the last key]) ----
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nj = 0; nj < NT; ++nj) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const bool ok = TAIL || 8 * nj + 2 * t + e < frames;
                s[nj][e] = ok ? s[nj][e] * scale_log2 : -INFINITY;
                s[nj][2 + e] = ok ? s[nj][2 + e] * scale_log2 : -INFINITY;
                mx0 = fmaxf(mx0, s[nj][e]);
                mx1 = fmaxf(mx1, s[nj][2 + e]);
            }
        }
        mx0 = fmaxf(quad_max(mx0), sl0);
        mx1 = fmaxf(quad_max(mx1), sl1);
        uint32_t pa[NT][2];            // fp16 P: [nj][0] = row r0, [nj][1] = row r1
        float sum0 = 0.f, sum1 = 0.f, sum0_32 = 0.f, sum1_32 = 0.f;
#pragma unroll
        for (int nj = 0; nj < NT; ++nj) {
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nj][e] = ex2f(s[nj][e] - (e < 2 ? mx0 : mx1));
            sum0_32 += s[nj][0] + s[nj][1];
            sum1_32 += s[nj][2] + s[nj][3];
            // the P.V MMA consumes fp16 P: normalise by the sum of the rounded values
            pa[nj][0] = pack_f16x2(s[nj][0], s[nj][1]);
            pa[nj][1] = pack_f16x2(s[nj][2], s[nj][3]);
            const float2 f0 = unpack_f16x2(pa[nj][0]), f1 = unpack_f16x2(pa[nj][1]);
            sum0 += f0.x + f0.y;
            sum1 += f1.x + f1.y;
        }
        const float pl0 = TAIL ? ex2f(sl0 - mx0) : 0.f, pl1 = TAIL ? ex2f(sl1 - mx1) : 0.f;
        sum0 = quad_sum(sum0) + pl0;
        sum1 = quad_sum(sum1) + pl1;
        const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;

        if (pr_base != nullptr) {    // probs[b, h, pos, i, j], fp32, from the unrounded exponentials
            const float i0 = 1.0f / (quad_sum(sum0_32) + pl0), i1 = 1.0f / (quad_sum(sum1_32) + pl1);
#pragma unroll
            for (int nj = 0; nj < NT; ++nj)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int key = 8 * nj + 2 * t + e;
                    if (key < fmain) {
                        if (r0 < fmain) pr_base[static_cast<int64_t>(r0) * frames + key] = s[nj][e] * i0;
                        if (r1 < fmain) pr_base[static_cast<int64_t>(r1) * frames + key] = s[nj][2 + e] * i1;
                    }
                }
            if (TAIL && t == 0) {
                pr_base[static_cast<int64_t>(r0) * frames + R] = pl0 * i0;
                pr_base[static_cast<int64_t>(r1) * frames + R] = pl1 * i1;
            }
        }

        // ---- O = P V: lane owns dims [16t, 16t+16) of rows r0 and r1 ----
        float o[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            if constexpr (TAIL) {        // the last key's contribution, fp32
                o[nt][0] = pl0 * vtf[nt]; o[nt][1] = pl0 * vtf[8 + nt];
                o[nt][2] = pl1 * vtf[nt]; o[nt][3] = pl1 * vtf[8 + nt];
            } else {
                o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
            }
        }
#pragma unroll
        for (int kk = 0; kk < MT; ++kk) {
            if (16 * kk >= fmain) break;               // warp-uniform
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
                mma_f16_16816(o[nt], pa[2 * kk][0], pa[2 * kk][1], pa[2 * kk + 1][0], pa[2 * kk + 1][1], vb[kk][nt][0],
                              vb[kk][nt][1]);
        }
        auto store_row = [&](int r, float inv, int e0) {
            __nv_bfloat16* op = out + (row0 + static_cast<int64_t>(r) * tokens) * inner + h * TA_DH + 16 * t;
            uint4 a, c;
            a.x = pack_bf16x2_rn(o[0][e0] * inv, o[1][e0] * inv); a.y = pack_bf16x2_rn(o[2][e0] * inv, o[3][e0] * inv);
            a.z = pack_bf16x2_rn(o[4][e0] * inv, o[5][e0] * inv); a.w = pack_bf16x2_rn(o[6][e0] * inv, o[7][e0] * inv);
            c.x = pack_bf16x2_rn(o[0][e0 + 1] * inv, o[1][e0 + 1] * inv); c.y = pack_bf16x2_rn(o[2][e0 + 1] * inv, o[3][e0 + 1] * inv);
            c.z = pack_bf16x2_rn(o[4][e0 + 1] * inv, o[5][e0 + 1] * inv); c.w = pack_bf16x2_rn(o[6][e0 + 1] * inv, o[7][e0 + 1] * inv);
            *reinterpret_cast<uint4*>(op) = a;
            *reinterpret_cast<uint4*>(op + 8) = c;
        };
        if (r0 < fmain) store_row(r0, inv0, 0);
        if (r1 < fmain) store_row(r1, inv1, 2);
        if (mi + 1 < MT) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) { qa[rr][0] = qn[rr][0]; qa[rr][1] = qn[rr][1]; }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Backward of the temporal attention for F <= 8 frames, same one-warp-per-(clip, position, head) mma.sync
// scheme as the forward (everything stays in registers; probabilities are recomputed):
//   S = Q K^T, P = softmax(S * scale), dP = dO V^T, D_i = sum_j P_ij dP_ij, dS = P o (dP - D) * scale
//   dV = P^T dO,  dQ = dS K,  dK = dS^T Q
// P^T and dS^T come from movmatrix (8x8 b16 transpose across the warp); the B operands whose k index is the
// frame ({X[2t][d], X[2t+1][d]}) are byte-permutes of two 16-byte row loads, as for V in the forward.
// dqk: [rows, 2*heads*64] (dQ columns then dK columns), dv: [rows, heads*64].
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}

__global__ void __launch_bounds__(256)
attn_temporal_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qk, const __nv_bfloat16* __restrict__ v,
                             const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqk,
                             __nv_bfloat16* __restrict__ dv, float* __restrict__ cam, int frames, int tokens,
                             int heads, float scale, int64_t units) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2;
    const int t = lane & 3;
    const int64_t unit = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (unit >= units) return;
    const int h = static_cast<int>(unit % heads);
    const int64_t bp = unit / heads;
    const int64_t b = bp / tokens;
    const int pos = static_cast<int>(bp - b * tokens);
    const int inner = heads * TA_DH;
    const int64_t row0 = b * frames * tokens + pos;
    const float scale_log2 = scale * 1.4426950408889634f;

    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    // pattern A: row g, 16-byte chunks t and t+4 of the head's 64 dims
    uint4 qa0 = zero, qa1 = zero, ka0 = zero, ka1 = zero, va0 = zero, va1 = zero, oa0 = zero, oa1 = zero;
    if (g < frames) {
        const int64_t r = row0 + static_cast<int64_t>(g) * tokens;
        const __nv_bfloat16* qp = qk + r * (2 * inner) + h * TA_DH;
        const __nv_bfloat16* vp = v + r * inner + h * TA_DH;
        const __nv_bfloat16* op = dout + r * inner + h * TA_DH;
        qa0 = ldg_nc_u4(qp + 8 * t);         qa1 = ldg_nc_u4(qp + 8 * (t + 4));
        ka0 = ldg_nc_u4(qp + inner + 8 * t); ka1 = ldg_nc_u4(qp + inner + 8 * (t + 4));
        va0 = ldg_nc_u4(vp + 8 * t);         va1 = ldg_nc_u4(vp + 8 * (t + 4));
        oa0 = ldg_nc_u4(op + 8 * t);         oa1 = ldg_nc_u4(op + 8 * (t + 4));
    }
    // pattern B: rows 2t and 2t+1, dims [8g, 8g+8)
    uint4 qb0 = zero, qb1 = zero, kb0 = zero, kb1 = zero, ob0 = zero, ob1 = zero;
    if (2 * t < frames) {
        const int64_t r = row0 + static_cast<int64_t>(2 * t) * tokens;
        qb0 = ldg_nc_u4(qk + r * (2 * inner) + h * TA_DH + 8 * g);
        kb0 = ldg_nc_u4(qk + r * (2 * inner) + inner + h * TA_DH + 8 * g);
        ob0 = ldg_nc_u4(dout + r * inner + h * TA_DH + 8 * g);
    }
    if (2 * t + 1 < frames) {
        const int64_t r = row0 + static_cast<int64_t>(2 * t + 1) * tokens;
        qb1 = ldg_nc_u4(qk + r * (2 * inner) + h * TA_DH + 8 * g);
        kb1 = ldg_nc_u4(qk + r * (2 * inner) + inner + h * TA_DH + 8 * g);
        ob1 = ldg_nc_u4(dout + r * inner + h * TA_DH + 8 * g);
    }

    float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
    mma_bf16_16816(s, qa0.x, 0u, qa0.y, 0u, ka0.x, ka0.y);
    mma_bf16_16816(s, qa0.z, 0u, qa0.w, 0u, ka0.z, ka0.w);
    mma_bf16_16816(s, qa1.x, 0u, qa1.y, 0u, ka1.x, ka1.y);
    mma_bf16_16816(s, qa1.z, 0u, qa1.w, 0u, ka1.z, ka1.w);
    mma_bf16_16816(dp, oa0.x, 0u, oa0.y, 0u, va0.x, va0.y);
    mma_bf16_16816(dp, oa0.z, 0u, oa0.w, 0u, va0.z, va0.w);
    mma_bf16_16816(dp, oa1.x, 0u, oa1.y, 0u, va1.x, va1.y);
    mma_bf16_16816(dp, oa1.z, 0u, oa1.w, 0u, va1.z, va1.w);

    const float s0 = (2 * t < frames) ? s[0] * scale_log2 : -INFINITY;
    const float s1 = (2 * t + 1 < frames) ? s[1] * scale_log2 : -INFINITY;
    float mx = fmaxf(s0, s1);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float p0, p1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(s0 - mx));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(s1 - mx));
    float sum = p0 + p1;
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
    p0 *= inv; p1 *= inv;
    float dsum = p0 * dp[0] + p1 * dp[1];
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
    const float ds0 = p0 * (dp[0] - dsum) * scale;
    const float ds1 = p1 * (dp[1] - dsum) * scale;
    if (cam != nullptr && g < frames) {
        // relevance pass: cam[b, pos, i, j] += relu(dA o A) / heads
        float* cr = cam + ((b * tokens + pos) * frames + g) * frames;
        const float ih = 1.0f / static_cast<float>(heads);
        if (2 * t < frames) atomicAdd(cr + 2 * t, fmaxf(p0 * dp[0], 0.f) * ih);
        if (2 * t + 1 < frames) atomicAdd(cr + 2 * t + 1, fmaxf(p1 * dp[1], 0.f) * ih);
    }
    auto pk = [](float x, float y) {
        const __nv_bfloat162 r = __floats2bfloat162_rn(x, y);
        return *reinterpret_cast<const uint32_t*>(&r);
    };
    const uint32_t pp = pk(p0, p1);
    const uint32_t dsp = pk(ds0, ds1);
    const uint32_t ppT = movmatrix_trans(pp);
    const uint32_t dsT = movmatrix_trans(dsp);

    const uint32_t wq0[4] = {qb0.x, qb0.y, qb0.z, qb0.w}, wq1[4] = {qb1.x, qb1.y, qb1.z, qb1.w};
    const uint32_t wk0[4] = {kb0.x, kb0.y, kb0.z, kb0.w}, wk1[4] = {kb1.x, kb1.y, kb1.z, kb1.w};
    const uint32_t wo0[4] = {ob0.x, ob0.y, ob0.z, ob0.w}, wo1[4] = {ob1.x, ob1.y, ob1.z, ob1.w};
    float dq_lo[8], dq_hi[8], dk_lo[8], dk_hi[8], dv_lo[8], dv_hi[8];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const uint32_t sel = (nt & 1) ? 0x7632 : 0x5410;
        float d[4];
        d[0] = d[1] = d[2] = d[3] = 0.f;
        mma_bf16_16816(d, ppT, 0u, 0u, 0u, __byte_perm(wo0[nt >> 1], wo1[nt >> 1], sel), 0u);   // dV = P^T dO
        dv_lo[nt] = d[0]; dv_hi[nt] = d[1];
        d[0] = d[1] = d[2] = d[3] = 0.f;
        mma_bf16_16816(d, dsp, 0u, 0u, 0u, __byte_perm(wk0[nt >> 1], wk1[nt >> 1], sel), 0u);   // dQ = dS K
        dq_lo[nt] = d[0]; dq_hi[nt] = d[1];
        d[0] = d[1] = d[2] = d[3] = 0.f;
        mma_bf16_16816(d, dsT, 0u, 0u, 0u, __byte_perm(wq0[nt >> 1], wq1[nt >> 1], sel), 0u);   // dK = dS^T Q
        dk_lo[nt] = d[0]; dk_hi[nt] = d[1];
    }
    if (g < frames) {
        const int64_t r = row0 + static_cast<int64_t>(g) * tokens;
        auto store16 = [&](__nv_bfloat16* dst, const float (&lo)[8], const float (&hi)[8]) {
            uint4 a, c;
            a.x = pk(lo[0], lo[1]); a.y = pk(lo[2], lo[3]); a.z = pk(lo[4], lo[5]); a.w = pk(lo[6], lo[7]);
            c.x = pk(hi[0], hi[1]); c.y = pk(hi[2], hi[3]); c.z = pk(hi[4], hi[5]); c.w = pk(hi[6], hi[7]);
            *reinterpret_cast<uint4*>(dst) = a;
            *reinterpret_cast<uint4*>(dst + 8) = c;
        };
        store16(dqk + r * (2 * inner) + h * TA_DH + 16 * t, dq_lo, dq_hi);
        store16(dqk + r * (2 * inner) + inner + h * TA_DH + 16 * t, dk_lo, dk_hi);
        store16(dv + r * inner + h * TA_DH + 16 * t, dv_lo, dv_hi);
    }
}

// ------------------------------------------------------------------------------------------
// Backward of the temporal attention for 8 < F <= 16 * MT frames (MT = 2, 3): one warp per (clip, position, head) on
// mma.sync again, but with the four operand matrices of the unit (Q, K, V, dO: RP = 16 MT rows x 64 dims, rows >= F
// zero-filled) staged ONCE in shared memory by cp.async (16-byte chunks XOR-swizzled by row) and read back with
// ldmatrix — every tensor is needed in two fragment roles here (K: B of Q.K^T with k = dim, and B of dS.K with
// k = key; Q: A of Q.K^T, and B of dS^T.Q; ...), which the registers-only scheme of the forward cannot hold.
//   phase A, per 16-query tile mi:  S = Q K^T, dP = dO V^T, P = softmax(S scale), D = rowsum(P o dP),
//                                   dS = P o (dP - D) scale, dQ[mi] = dS K; P and dS stay in registers as bf16 pairs
//   phase B, per 16-key tile kj:    dV[kj] = sum_mi P[mi, kj]^T dO[mi], dK[kj] = sum_mi dS[mi, kj]^T Q[mi]
//                                   (A fragments = movmatrix transposes of the 8x8 blocks of P / dS)
// Zero rows are inert: a padded query row has dO = 0, hence dP = D = dS = 0 and no contribution to dV / dK.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// byte offset of 16-byte chunk `chunk` of row `row` in a [rows][64] bf16 tile (128-byte rows)
__device__ __forceinline__ uint32_t tw_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

template <int MT>
__global__ void __launch_bounds__(128)
attn_temporal_bwd_wide_kernel(const __nv_bfloat16* __restrict__ qk, const __nv_bfloat16* __restrict__ v,
                              const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqk,
                              __nv_bfloat16* __restrict__ dv, float* __restrict__ cam, int frames, int tokens,
                              int heads, float scale, int64_t units) {
    constexpr int NT = 2 * MT;
    constexpr int RP = 16 * MT;                      // padded rows per tensor
    constexpr int TILE_BYTES = RP * 128;
    extern __shared__ __align__(128) uint8_t tw_smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2;
    const int t = lane & 3;
    const int64_t unit = static_cast<int64_t>(blockIdx.x) * 4 + warp;
    if (unit >= units) return;
    const int h = static_cast<int>(unit % heads);
    const int64_t bp = unit / heads;
    const int64_t b = bp / tokens;
    const int pos = static_cast<int>(bp - b * tokens);
    const int inner = heads * TA_DH;
    const int64_t row0 = b * frames * tokens + pos;
    const float scale_log2 = scale * 1.4426950408889634f;

    const uint32_t sq = smem_u32(tw_smem) + warp * (4 * TILE_BYTES);
    const uint32_t sk = sq + TILE_BYTES, sv = sk + TILE_BYTES, sdo = sv + TILE_BYTES;

    // ---- stage Q, K, V, dO (zero-fill rows >= frames) ----
    {
        const __nv_bfloat16* qb = qk + row0 * (2 * inner) + h * TA_DH;
        const __nv_bfloat16* vb = v + row0 * inner + h * TA_DH;
        const __nv_bfloat16* ob = dout + row0 * inner + h * TA_DH;
        for (int i = lane; i < RP * 8; i += 32) {
            const int row = i >> 3, chunk = i & 7;
            const bool ok = row < frames;
            const int64_t r = ok ? row : 0;
            const uint32_t n = ok ? 16u : 0u;
            const uint32_t off = tw_off(row, chunk);
            const __nv_bfloat16* qp = qb + r * tokens * (2 * inner) + chunk * 8;
            const __nv_bfloat16* vp = vb + r * tokens * inner + chunk * 8;
            const __nv_bfloat16* op = ob + r * tokens * inner + chunk * 8;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sq + off), "l"(qp), "r"(n) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sk + off), "l"(qp + inner), "r"(n) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sv + off), "l"(vp), "r"(n) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdo + off), "l"(op), "r"(n) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
    }

    auto pk = [](float x, float y) {
        const __nv_bfloat162 r = __floats2bfloat162_rn(x, y);
        return *reinterpret_cast<const uint32_t*>(&r);
    };
    // ldmatrix lane addressing (see the fragment maps in the kernel comment)
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_chunk = lane >> 4;     // A tiles and transposed B tiles
    const int b_row = lane & 7, b_chunk = lane >> 3;                               // B tiles of X.Y^T (k = head dim)

    uint32_t pp[MT][NT][2], dsp[MT][NT][2];       // bf16 pairs: [..][0] = row 16 mi + g, [..][1] = row 16 mi + g + 8

    // ================= phase A =================
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
        uint32_t qa[4][4], da[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            ldsm_x4(sq + tw_off(16 * mi + a_row, 2 * ks + a_chunk), qa[ks]);
            ldsm_x4(sdo + tw_off(16 * mi + a_row, 2 * ks + a_chunk), da[ks]);
        }
        float s[NT][4], dp[NT][4];
#pragma unroll
        for (int nj = 0; nj < NT; ++nj) {
            s[nj][0] = s[nj][1] = s[nj][2] = s[nj][3] = 0.f;
            dp[nj][0] = dp[nj][1] = dp[nj][2] = dp[nj][3] = 0.f;
#pragma unroll
            for (int hk = 0; hk < 2; ++hk) {          // 32 head dims per ldmatrix.x4
                uint32_t kb[4], vb4[4];
                ldsm_x4(sk + tw_off(8 * nj + b_row, 4 * hk + b_chunk), kb);
                ldsm_x4(sv + tw_off(8 * nj + b_row, 4 * hk + b_chunk), vb4);
                mma_bf16_16816(s[nj], qa[2 * hk][0], qa[2 * hk][1], qa[2 * hk][2], qa[2 * hk][3], kb[0], kb[1]);
                mma_bf16_16816(s[nj], qa[2 * hk + 1][0], qa[2 * hk + 1][1], qa[2 * hk + 1][2], qa[2 * hk + 1][3], kb[2], kb[3]);
                mma_bf16_16816(dp[nj], da[2 * hk][0], da[2 * hk][1], da[2 * hk][2], da[2 * hk][3], vb4[0], vb4[1]);
                mma_bf16_16816(dp[nj], da[2 * hk + 1][0], da[2 * hk + 1][1], da[2 * hk + 1][2], da[2 * hk + 1][3], vb4[2], vb4[3]);
            }
        }
        // ---- softmax rows r0 = 16 mi + g (elements 0, 1) and r1 = r0 + 8 (elements 2, 3) ----
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nj = 0; nj < NT; ++nj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const bool ok = 8 * nj + 2 * t + e < frames;
                s[nj][e] = ok ? s[nj][e] * scale_log2 : -INFINITY;
                s[nj][2 + e] = ok ? s[nj][2 + e] * scale_log2 : -INFINITY;
                mx0 = fmaxf(mx0, s[nj][e]);
                mx1 = fmaxf(mx1, s[nj][2 + e]);
            }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int nj = 0; nj < NT; ++nj)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float pe;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe) : "f"(s[nj][e] - (e < 2 ? mx0 : mx1)));
                s[nj][e] = pe;
                if (e < 2) sum0 += pe; else sum1 += pe;
            }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int nj = 0; nj < NT; ++nj) {
            s[nj][0] *= inv0; s[nj][1] *= inv0; s[nj][2] *= inv1; s[nj][3] *= inv1;
            d0 += s[nj][0] * dp[nj][0] + s[nj][1] * dp[nj][1];
            d1 += s[nj][2] * dp[nj][2] + s[nj][3] * dp[nj][3];
        }
        d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
        d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
        d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
        d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
        const int r0 = 16 * mi + g, r1 = r0 + 8;
        if (cam != nullptr) {
            // relevance pass: cam[b, pos, i, j] += relu(dA o A) / heads
            float* cr = cam + (b * tokens + pos) * frames * frames;
            const float ih = 1.0f / static_cast<float>(heads);
#pragma unroll
            for (int nj = 0; nj < NT; ++nj)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int key = 8 * nj + 2 * t + e;
                    if (key < frames) {
                        if (r0 < frames) atomicAdd(cr + r0 * frames + key, fmaxf(s[nj][e] * dp[nj][e], 0.f) * ih);
                        if (r1 < frames) atomicAdd(cr + r1 * frames + key, fmaxf(s[nj][2 + e] * dp[nj][2 + e], 0.f) * ih);
                    }
                }
        }
#pragma unroll
        for (int nj = 0; nj < NT; ++nj) {
            pp[mi][nj][0] = pk(s[nj][0], s[nj][1]);
            pp[mi][nj][1] = pk(s[nj][2], s[nj][3]);
            dsp[mi][nj][0] = pk(s[nj][0] * (dp[nj][0] - d0) * scale, s[nj][1] * (dp[nj][1] - d0) * scale);
            dsp[mi][nj][1] = pk(s[nj][2] * (dp[nj][2] - d1) * scale, s[nj][3] * (dp[nj][3] - d1) * scale);
        }
        // ---- dQ[mi] = dS[mi] K ----
        float dq[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < MT; ++kk)
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {           // 16 head dims per ldmatrix.x4.trans
                uint32_t kb[4];
                ldsm_x4_trans(sk + tw_off(16 * kk + a_row, 2 * n2 + a_chunk), kb);
                mma_bf16_16816(dq[2 * n2], dsp[mi][2 * kk][0], dsp[mi][2 * kk][1], dsp[mi][2 * kk + 1][0],
                               dsp[mi][2 * kk + 1][1], kb[0], kb[1]);
                mma_bf16_16816(dq[2 * n2 + 1], dsp[mi][2 * kk][0], dsp[mi][2 * kk][1], dsp[mi][2 * kk + 1][0],
                               dsp[mi][2 * kk + 1][1], kb[2], kb[3]);
            }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            if (r0 < frames)
                *reinterpret_cast<uint32_t*>(dqk + (row0 + static_cast<int64_t>(r0) * tokens) * (2 * inner) + h * TA_DH + 8 * nt + 2 * t) =
                    pk(dq[nt][0], dq[nt][1]);
            if (r1 < frames)
                *reinterpret_cast<uint32_t*>(dqk + (row0 + static_cast<int64_t>(r1) * tokens) * (2 * inner) + h * TA_DH + 8 * nt + 2 * t) =
                    pk(dq[nt][2], dq[nt][3]);
        }
    }

    // ================= phase B =================
#pragma unroll
    for (int kj = 0; kj < MT; ++kj) {
        if (16 * kj >= frames) break;          // warp-uniform
        float dvv[8][4], dkk[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            dvv[nt][0] = dvv[nt][1] = dvv[nt][2] = dvv[nt][3] = 0.f;
            dkk[nt][0] = dkk[nt][1] = dkk[nt][2] = dkk[nt][3] = 0.f;
        }
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
            const uint32_t pa0 = movmatrix_trans(pp[mi][2 * kj][0]), pa1 = movmatrix_trans(pp[mi][2 * kj + 1][0]);
            const uint32_t pa2 = movmatrix_trans(pp[mi][2 * kj][1]), pa3 = movmatrix_trans(pp[mi][2 * kj + 1][1]);
            const uint32_t sa0 = movmatrix_trans(dsp[mi][2 * kj][0]), sa1 = movmatrix_trans(dsp[mi][2 * kj + 1][0]);
            const uint32_t sa2 = movmatrix_trans(dsp[mi][2 * kj][1]), sa3 = movmatrix_trans(dsp[mi][2 * kj + 1][1]);
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                uint32_t ob[4], qb4[4];
                ldsm_x4_trans(sdo + tw_off(16 * mi + a_row, 2 * n2 + a_chunk), ob);
                ldsm_x4_trans(sq + tw_off(16 * mi + a_row, 2 * n2 + a_chunk), qb4);
                mma_bf16_16816(dvv[2 * n2], pa0, pa1, pa2, pa3, ob[0], ob[1]);
                mma_bf16_16816(dvv[2 * n2 + 1], pa0, pa1, pa2, pa3, ob[2], ob[3]);
                mma_bf16_16816(dkk[2 * n2], sa0, sa1, sa2, sa3, qb4[0], qb4[1]);
                mma_bf16_16816(dkk[2 * n2 + 1], sa0, sa1, sa2, sa3, qb4[2], qb4[3]);
            }
        }
        const int r0 = 16 * kj + g, r1 = r0 + 8;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            if (r0 < frames) {
                const int64_t r = row0 + static_cast<int64_t>(r0) * tokens;
                *reinterpret_cast<uint32_t*>(dqk + r * (2 * inner) + inner + h * TA_DH + 8 * nt + 2 * t) = pk(dkk[nt][0], dkk[nt][1]);
                *reinterpret_cast<uint32_t*>(dv + r * inner + h * TA_DH + 8 * nt + 2 * t) = pk(dvv[nt][0], dvv[nt][1]);
            }
            if (r1 < frames) {
                const int64_t r = row0 + static_cast<int64_t>(r1) * tokens;
                *reinterpret_cast<uint32_t*>(dqk + r * (2 * inner) + inner + h * TA_DH + 8 * nt + 2 * t) = pk(dkk[nt][2], dkk[nt][3]);
                *reinterpret_cast<uint32_t*>(dv + r * inner + h * TA_DH + 8 * nt + 2 * t) = pk(dvv[nt][2], dvv[nt][3]);
            }
        }
    }
}

template <typename T>
static int launch_temporal(const void* qk, const void* v, void* out, float* probs, int batch, int frames,
                           int tokens, int heads, float scale, cudaStream_t st) {
    const int threads = ((heads * frames + 31) / 32) * 32;
    ISTVT_REQUIRE(threads <= 288);
    const size_t smem = 2 * static_cast<size_t>(frames) * heads * TA_ROW * sizeof(T);
    auto kern = attn_temporal_kernel<T>;
    if (smem > 48 * 1024)
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<batch * tokens, threads, smem, st>>>(static_cast<const T*>(qk), static_cast<const T*>(v),
                                               static_cast<T*>(out), probs, frames, tokens, heads, scale);
    count_launch();
    return launch_status();
}

}  // namespace istvt

using namespace istvt;

extern "C" int istvt_attn_temporal_fwd(const void* qk, const void* v, void* out, float* probs, int dtype, int batch,
                                       int frames, int tokens, int heads, float scale, istvt_stream_t stream) {
    ISTVT_REQUIRE(qk && v && out);
    ISTVT_REQUIRE(batch > 0 && frames > 0 && tokens > 0 && heads > 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == ISTVT_BF16 && frames <= 8) {
        ISTVT_REQUIRE(((reinterpret_cast<uintptr_t>(qk) | reinterpret_cast<uintptr_t>(v) |
                        reinterpret_cast<uintptr_t>(out)) & 15) == 0);
        const int64_t units = static_cast<int64_t>(batch) * tokens * heads;
        const int64_t grid = (units + 7) / 8;
        ISTVT_REQUIRE(grid < (int64_t(1) << 31));
        attn_temporal_mma_kernel<<<static_cast<unsigned>(grid), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(qk), static_cast<const __nv_bfloat16*>(v),
            static_cast<__nv_bfloat16*>(out), probs, frames, tokens, heads, scale * 1.4426950408889634f, units);
        count_launch();
        return launch_status();
    }
    if (dtype == ISTVT_BF16 && frames <= 48) {
        ISTVT_REQUIRE(((reinterpret_cast<uintptr_t>(qk) | reinterpret_cast<uintptr_t>(v) |
                        reinterpret_cast<uintptr_t>(out)) & 15) == 0);
        const int64_t units = static_cast<int64_t>(batch) * tokens * heads;
        const int64_t grid = (units + 3) / 4;
        ISTVT_REQUIRE(grid < (int64_t(1) << 31));
        const float sl2 = scale * 1.4426950408889634f;
        const __nv_bfloat16* qkp = static_cast<const __nv_bfloat16*>(qk);
        const __nv_bfloat16* vp = static_cast<const __nv_bfloat16*>(v);
        __nv_bfloat16* op = static_cast<__nv_bfloat16*>(out);
        const unsigned gr = static_cast<unsigned>(grid);
        // F = 16 MT + 1 (T = 16 / 32 frames + the temporal class frame): core tiles on mma.sync, last frame on the FMA pipe
        static const bool tail_env = []() { const char* e = getenv("ISTVT_TA_TAIL"); return !e || atoi(e) != 0; }();
        if (frames == 33 && tail_env)
            attn_temporal_mma_wide_kernel<2, true><<<gr, 128, 0, st>>>(qkp, vp, op, probs, frames, tokens, heads, sl2, units);
        else if (frames == 17 && tail_env)
            attn_temporal_mma_wide_kernel<1, true><<<gr, 128, 0, st>>>(qkp, vp, op, probs, frames, tokens, heads, sl2, units);
        else if (frames <= 16)
            attn_temporal_mma_wide_kernel<1, false><<<gr, 128, 0, st>>>(qkp, vp, op, probs, frames, tokens, heads, sl2, units);
        else if (frames <= 32)
            attn_temporal_mma_wide_kernel<2, false><<<gr, 128, 0, st>>>(qkp, vp, op, probs, frames, tokens, heads, sl2, units);
        else
            attn_temporal_mma_wide_kernel<3, false><<<gr, 128, 0, st>>>(qkp, vp, op, probs, frames, tokens, heads, sl2, units);
        count_launch();
        return launch_status();
    }
    if (heads * frames > 288) return ISTVT_ERR_UNSUPPORTED;    // SIMT kernel (fp32 validation mode): F <= 36
    ISTVT_REQUIRE(static_cast<int64_t>(batch) * tokens < (int64_t(1) << 31));
    if (dtype == ISTVT_BF16)
        return launch_temporal<__nv_bfloat16>(qk, v, out, probs, batch, frames, tokens, heads, scale, st);
    if (dtype == ISTVT_F32) return launch_temporal<float>(qk, v, out, probs, batch, frames, tokens, heads, scale, st);
    return ISTVT_ERR_INVALID_ARG;
}

template <int MT>
static int launch_temporal_bwd_wide(const void* qk, const void* v, const void* dout, void* dqk, void* dv, float* cam,
                                    int frames, int tokens, int heads, float scale, int64_t units,
                                    cudaStream_t st) {
    const int smem = 4 * 4 * (16 * MT) * 128;      // 4 warps x {Q, K, V, dO} x padded rows x 128 B
    auto kern = attn_temporal_bwd_wide_kernel<MT>;
    ISTVT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t grid = (units + 3) / 4;
    ISTVT_REQUIRE(grid < (int64_t(1) << 31));
    kern<<<static_cast<unsigned>(grid), 128, smem, st>>>(
        static_cast<const __nv_bfloat16*>(qk), static_cast<const __nv_bfloat16*>(v),
        static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(dqk), static_cast<__nv_bfloat16*>(dv), cam,
        frames, tokens, heads, scale, units);
    count_launch();
    return launch_status();
}

static int attn_temporal_bwd_launch(const void* qk, const void* v, const void* dout, void* dqk, void* dv, float* cam,
                                    int batch, int frames, int tokens, int heads, float scale,
                                    istvt_stream_t stream) {
    ISTVT_REQUIRE(qk && v && dout && dqk && dv);
    ISTVT_REQUIRE(batch > 0 && frames > 0 && tokens > 0 && heads > 0);
    if (frames > 48) return ISTVT_ERR_UNSUPPORTED;   // mma.sync kernels cover T <= 47 frames (C5: T = 32)
    ISTVT_REQUIRE(((reinterpret_cast<uintptr_t>(qk) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(dout) |
                    reinterpret_cast<uintptr_t>(dqk) | reinterpret_cast<uintptr_t>(dv)) & 15) == 0);
    const int64_t units = static_cast<int64_t>(batch) * tokens * heads;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (frames > 32)
        return launch_temporal_bwd_wide<3>(qk, v, dout, dqk, dv, cam, frames, tokens, heads, scale, units, st);
    if (frames > 8)
        return launch_temporal_bwd_wide<2>(qk, v, dout, dqk, dv, cam, frames, tokens, heads, scale, units, st);
    const int64_t grid = (units + 7) / 8;
    ISTVT_REQUIRE(grid < (int64_t(1) << 31));
    attn_temporal_bwd_mma_kernel<<<static_cast<unsigned>(grid), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(qk), static_cast<const __nv_bfloat16*>(v),
        static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(dqk), static_cast<__nv_bfloat16*>(dv), cam,
        frames, tokens, heads, scale, units);
    count_launch();
    return launch_status();
}

extern "C" int istvt_attn_temporal_bwd(const void* qk, const void* v, const void* dout, void* dqk, void* dv, int batch,
                                       int frames, int tokens, int heads, float scale, istvt_stream_t stream) {
    return attn_temporal_bwd_launch(qk, v, dout, dqk, dv, nullptr, batch, frames, tokens, heads, scale, stream);
}

// Relevance pass variant: cam[batch, tokens, frames, frames] (fp32, zero-filled by the caller) += relu(dA o A) / heads.
extern "C" int istvt_attn_temporal_bwd_cam(const void* qk, const void* v, const void* dout, void* dqk, void* dv, float* cam,
                                           int batch, int frames, int tokens, int heads, float scale,
                                           istvt_stream_t stream) {
    ISTVT_REQUIRE(cam != nullptr);
    return attn_temporal_bwd_launch(qk, v, dout, dqk, dv, cam, batch, frames, tokens, heads, scale, stream);
}
