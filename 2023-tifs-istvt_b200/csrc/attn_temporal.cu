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
#include "simt_util.cuh"

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
                             __nv_bfloat16* __restrict__ dv, float* __restrict__ cam, int frames, int tokens, int heads,
                             float scale, int64_t units) {
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
    ISTVT_REQUIRE(heads * frames <= 288);
    ISTVT_REQUIRE(static_cast<int64_t>(batch) * tokens < (int64_t(1) << 31));
    if (dtype == ISTVT_BF16)
        return launch_temporal<__nv_bfloat16>(qk, v, out, probs, batch, frames, tokens, heads, scale, st);
    if (dtype == ISTVT_F32) return launch_temporal<float>(qk, v, out, probs, batch, frames, tokens, heads, scale, st);
    return ISTVT_ERR_INVALID_ARG;
}

static int attn_temporal_bwd_launch(const void* qk, const void* v, const void* dout, void* dqk, void* dv, float* cam,
                                    int batch, int frames, int tokens, int heads, float scale, istvt_stream_t stream) {
    ISTVT_REQUIRE(qk && v && dout && dqk && dv);
    ISTVT_REQUIRE(batch > 0 && frames > 0 && tokens > 0 && heads > 0);
    if (frames > 8) return ISTVT_ERR_UNSUPPORTED;   // training is built for the ISTVT configuration (T = 6, F = 7)
    ISTVT_REQUIRE(((reinterpret_cast<uintptr_t>(qk) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(dout) |
                    reinterpret_cast<uintptr_t>(dqk) | reinterpret_cast<uintptr_t>(dv)) & 15) == 0);
    const int64_t units = static_cast<int64_t>(batch) * tokens * heads;
    const int64_t grid = (units + 7) / 8;
    ISTVT_REQUIRE(grid < (int64_t(1) << 31));
    attn_temporal_bwd_mma_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
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
