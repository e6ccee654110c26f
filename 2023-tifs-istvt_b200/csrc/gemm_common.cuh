// Shared pieces of the two tcgen05 GEMM kernels (gemm_tcgen05.cu: 1-CTA tiles + implicit-GEMM conv mode;
// gemm_tcgen05_2cta.cu: CTA-pair tiles): problem descriptor and the fused epilogue
//   C = act(acc + bias) + residual   ->  bf16 / fp32, 16-byte stores.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace istvt {

constexpr int GEMM_BLOCK_M = 128;   // rows per CTA (= TMEM lanes)
constexpr int UMMA_K = 16;          // bf16

struct GemmParams {
    int64_t M;        // GEMM rows (conv mode: n_img * h_in * w_in, the padded grid)
    int N, K;         // K = per-tap K in conv mode
    int taps;         // 1 (plain GEMM) or 9 (3x3 conv)
    int conv_w_in;    // conv mode: input width  (row shift of tap = ky * conv_w_in + kx)
    int conv_h_in;    // conv mode: input height
    int pair;         // conv mode on PIXEL PAIRS (conv3x3_pair_launch): A rows are 2 adjacent pixels x cin, N = 2 x cout
    int tap_shift[9]; // conv mode: A row shift of every tap relative to the tile's first row
    int b_resident;   // 1-CTA kernel, one N tile: the whole weight matrix is loaded into shared memory once per CTA
    void* C;
    int64_t ldc;
    const float* bias;
    const float* residual;
    int64_t ldr;
    int act;
    int c_f32;
    int split_k;      // > 1 (CTA-pair kernel only): K is cut into split_k ranges of kb_per_split 64-wide k-blocks,
    int kb_per_split; // every (tile, range) adds its partial product to the fp32 C with red.global (C pre-initialised)
    int mn_major;     // CTA-pair kernel: both operands are stored K-rows x MN-contiguous (A = dY [k, m], B = X [k, n]):
                      // the weight-gradient GEMM dW = dY^T X reads dY and X as they sit in HBM, no transposed copies
#ifdef ISTVT_GEMM_TRACE
    int trace_no_tma; // debug builds only (tools/gemm_trace.py)
    int trace_no_epi;
#endif
    int split_producer;  // 1-CTA kernel: warp 0 loads the A boxes and warp 3 the B boxes (two issuing threads)
    int epi_tma;         // CTA-pair kernel, bf16 output: the epilogue's global stores are TMA box stores from the slabs
    // LayerNorm folded around a GEMM pair (CTA-pair kernel, bf16 output, TMA-store epilogue; engine.py, LN2).
    // Row statistics travel as one (sum, M2) pair per row and 64-column group — M2 = sum of squared deviations from the
    // GROUP mean — and are combined with Chan's formula: no atomics (bitwise reproducible), no E[x^2] - mu^2 cancellation.
    // Training-step fusions of the MLP (CTA-pair kernel, bf16 output, TMA-store epilogue):
    void* C2;                // ff1: second output, the PRE-activation acc + bias (bf16 [M, N], pitch ldc2) next to
    int64_t ldc2;            //      C = act(acc + bias): replaces the stand-alone GELU pass of the training forward
    const void* mul_pre;     // ff2 data gradient: C = acc * gelu'(mul_pre[m, n]) (bf16 [M, N], pitch ld_pre): replaces
    int64_t ld_pre;          //      the stand-alone GELU backward pass
    float2* row_stats_out;   // producer: fp32 [M, ceil(N / 64), 2], every slot written exactly once
    const float2* ln_stats;  // consumer: (mu, rstd) of ITS A rows, fp32 [M, 2];  C = rstd_m (acc - mu_m ln_c[n]) + ln_d[n]
    const float* ln_c;       //   fp32 [N]: row sums of the gamma-folded bf16 weight
    const float* ln_d;       //   fp32 [N]: W beta
};

// ------------------------------------------------------------------------------------------
// Fused epilogue, coalesced.  tcgen05.ld hands every thread ONE output row (TMEM lane) x 32 columns; storing
// that directly makes each warp-wide 16-byte access touch 32 different 128-byte lines (profiles/r1b: the
// L1->L2 request path was 73 % busy with half-filled sectors, and the fp32 read-modify-write of the residual
// stream ran at 255 TFLOP/s).  The smem data pipe (128 B/clk) is shared by the tensor core's operand reads,
// the TMA fills and every LSU wavefront, so the epilogue is written to minimise wavefronts:
// each epilogue warp owns 32 rows x 64 columns of the tile and transposes them through a private 4 KB slab
// (16-byte chunks XOR-swizzled by row, conflict-free both ways) so that afterwards 8 lanes cover 128
// contiguous bytes of one row:
//   bf16 output (no residual): bias + activation on the row-owner side, PACKED bf16 staged (64 B less per row),
//                              then one 16-byte store per lane = 4 full lines per warp instruction;
//   fp32 output (+ residual) : raw accumulators staged 32 columns at a time; bias / residual / store on full lines.
// ------------------------------------------------------------------------------------------
constexpr int EPI_SLAB_BYTES = 32 * 128;   // per epilogue warp
constexpr int EPI_SLAB_PLAIN_BYTES = 32 * 64;   // what the bf16-output path uses of it
constexpr int EPI_COLS = 64;               // accumulator columns per epilogue warp

// Output rows handled by this lane in the transposed phase: row (i*4 + lane/8) of the warp's 32, i = 0..7.
// `drow_lane` = destination row of the accumulator row this lane owns (TMEM lane), or -1 if it is not stored.
__device__ __forceinline__ void epilogue_rows(int drow_lane, int lane, int (&drow_t)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) drow_t[i] = __shfl_sync(0xffffffffu, drow_lane, i * 4 + (lane >> 3));
}

// Residual GEMMs (s_out, ff2) read-modify-write the fp32 token stream; those loads were the epilogue's critical
// path (DRAM latency, 4 x 16 B in flight per lane).  The addresses do not depend on the accumulator, so each
// epilogue warp pulls its rows' residual lines into L2 before it waits for the MMAs of the tile to finish.
__device__ __forceinline__ void epilogue_prefetch_residual(const GemmParams& p, int drow_lane, int n_first, int ncols) {
    if (p.residual == nullptr || drow_lane < 0) return;
    const float* row = p.residual + static_cast<int64_t>(drow_lane) * p.ldr;
    for (int n = n_first; n < n_first + ncols && n < p.N; n += 32)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(row + n));
}

__device__ __forceinline__ float epi_act(float v, int act) {
    if (act == ISTVT_ACT_RELU) return fmaxf(v, 0.f);
    if (act == ISTVT_ACT_GELU) return gelu_erf_fast(v);
    return v;
}

// Warp-collective epilogue of 32 rows x 64 columns [n0, n0+64) read from TMEM at `taddr` (lane quadrant and
// column already applied).  `after_tmem_reads()` is invoked once all TMEM reads of the call have completed.
// PLAIN_BF16 = bf16 output without residual (compile-time: keeps each kernel instantiation small — the first
// version inlined both paths twice and the 147 KB of SASS missed in the instruction cache).
// `slab` is the warp's staging area as a shared-space address; `drow_lane` = destination row of the accumulator row
// this lane owns (-1: not stored), `drow_t` = the same for the rows of the 8-lanes-per-row transposed phase.
// HALVES = 1 (bf16 path only): the warp owns 32 columns instead of 64 — twice the epilogue warps per tile for the
// short-K GEMMs whose tile period is the epilogue's latency chain.
template <bool PLAIN_BF16, int HALVES = 2, typename F>
__device__ __forceinline__ void gemm_epilogue_64(const GemmParams& p, uint32_t taddr, uint32_t slab, int drow_lane,
                                                 const int (&drow_t)[8], int n0, int lane, F after_tmem_reads) {
    const int cchunk = lane & 7;
    const int sw = lane & 7;
    if (n0 >= p.N) {   // whole column group is past the matrix edge (ragged last N tile): nothing to read or store
        after_tmem_reads();
        return;
    }
    if constexpr (PLAIN_BF16) {
        // ---------------- bf16 output: math first, packed staging, 32 columns (64 B per row) at a time ----------------
        // Only EPI_SLAB_PLAIN_BYTES (2 KB) of the slab are used: the CTA-pair kernel spends the other half of the
        // former 4 KB slabs on a sixth ring stage (its TMA ring is latency bound, profiles/README.md r3e).
        // 16-byte chunk q of row r sits at r*64 + ((q ^ ((r >> 1) & 3)) << 4): conflict-free for the row-owner writes
        // (8 lanes per phase, 64-byte pitch) and for the transposed reads (4 lanes per row).
        const uint32_t wrow = slab + lane * 64;
        const int wsw = (lane >> 1) & 3;
        const int cchunk4 = lane & 3;
#pragma unroll
        for (int hh = 0; hh < HALVES; ++hh) {
            const int nb = n0 + hh * 32;
            if (nb >= p.N) {   // warp-uniform: ragged last N tile
                if (hh == HALVES - 1) after_tmem_reads();
                break;
            }
            // the bias loads are issued BEFORE the TMEM load so that their latency hides behind it: behind it they sat
            // on the epilogue's critical path once per 8 columns (14 % of the samples of the conv2 kernel, whose tile
            // period is this epilogue: profiles/README.md r8l)
            float4 bb[8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int n = nb + g * 8;
                if (p.bias != nullptr && n < p.N) {
                    bb[2 * g] = ldg_nc_f4_ordered(p.bias + n);
                    bb[2 * g + 1] = ldg_nc_f4_ordered(p.bias + n + 4);
                } else {
                    bb[2 * g] = make_float4(0.f, 0.f, 0.f, 0.f);
                    bb[2 * g + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr + hh * 32, r);
            tmem_ld_wait();
            if (hh == HALVES - 1) after_tmem_reads();
#pragma unroll
            for (int g = 0; g < 4; ++g) {     // 8 columns -> one 16-byte chunk
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
                {
                    const float4 b0 = bb[2 * g], b1 = bb[2 * g + 1];
                    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = epi_act(v[j], p.act);
                sts_u4(wrow + ((g ^ wsw) << 4), pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                       pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            }
            __syncwarp();
            const int n = nb + cchunk4 * 8;
            const bool n_ok = n < p.N;                 // lane-dependent on the ragged last N tile: shuffles stay outside
            __nv_bfloat16* cbase = reinterpret_cast<__nv_bfloat16*>(p.C) + n;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = i * 8 + (lane >> 2);
                const int drow = __shfl_sync(0xffffffffu, drow_lane, row);
                if (n_ok && drow >= 0) {
                    const uint4 v = lds_u4(slab + row * 64 + ((cchunk4 ^ ((row >> 1) & 3)) << 4));
                    *reinterpret_cast<uint4*>(cbase + static_cast<int64_t>(drow) * p.ldc) = v;
                }
            }
            __syncwarp();   // the slab is rewritten by the next half
        }
    } else {
        // ---------------- fp32 output and/or residual: raw staging, 32 columns per pass ----------------
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int nb = n0 + hh * 32;
            uint32_t r[32];
            if (nb < p.N) {
                tmem_ld_32x32b_x32(taddr + hh * 32, r);
                tmem_ld_wait();
            }
            if (hh == 1) after_tmem_reads();
            if (nb >= p.N) break;
            const uint32_t wrow = slab + lane * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) sts_u4(wrow + ((c ^ sw) << 4), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
            __syncwarp();
            const int n = nb + cchunk * 4;
            const bool n_ok = n < p.N;
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n_ok && p.bias != nullptr) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll
            for (int i0 = 0; i0 < 8; i0 += 4) {
                float4 res[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.residual != nullptr && n_ok && drow_t[i0 + i] >= 0)
                        res[i] = *reinterpret_cast<const float4*>(p.residual + static_cast<int64_t>(drow_t[i0 + i]) * p.ldr + n);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = (i0 + i) * 4 + (lane >> 3);
                    const uint4 u = lds_u4(slab + row * 128 + ((cchunk ^ (row & 7)) << 4));
                    const int64_t drow = drow_t[i0 + i];
                    if (!n_ok || drow < 0) continue;
                    if (p.split_k > 1) {   // split-K partial: accumulate into C (weight-gradient GEMMs)
                        float* dst = reinterpret_cast<float*>(p.C) + drow * p.ldc + n;
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(u.x)),
                                     "f"(__uint_as_float(u.y)), "f"(__uint_as_float(u.z)), "f"(__uint_as_float(u.w))
                                     : "memory");
                        continue;
                    }
                    float4 v;
                    v.x = epi_act(__uint_as_float(u.x) + bias4.x, p.act) + res[i].x;
                    v.y = epi_act(__uint_as_float(u.y) + bias4.y, p.act) + res[i].y;
                    v.z = epi_act(__uint_as_float(u.z) + bias4.z, p.act) + res[i].z;
                    v.w = epi_act(__uint_as_float(u.w) + bias4.w, p.act) + res[i].w;
                    if (p.c_f32) {
                        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + drow * p.ldc + n) = v;
                    } else {
                        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.C) + drow * p.ldc + n) =
                            make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
                    }
                }
            }
            __syncwarp();   // slab is rewritten by the next pass
        }
    }
}

// bf16 output stored by TMA: the warp stages act(acc + bias) as packed bf16, 32 rows x 32 columns (64-byte rows) at a
// time, in its 2 KB slab — the row-owner writes of the transpose path above, whose chunk swizzle (q ^ ((r >> 1) & 3)) IS
// the TMA SWIZZLE_64B pattern for a 512-byte-aligned slab — and one lane issues cp.async.bulk.tensor of the box: no
// LDS / STG / address arithmetic per lane, ragged M / N edges clipped by the tensor map.  (Storing straight from
// registers with 32-byte vectors, one output row per lane, measured SLOWER: ff1 0.74 -> 0.79-0.89 ms, to_v 0.132 ->
// 0.154 ms, profiles/README.md r6d — 32 scattered sectors per store instruction cost more than the staging.)
// `row0` = first of the warp's 32 consecutive output rows.
// BOXC = 64: both 32-column halves are staged in one 4 KB slab (128-byte rows, SWIZZLE_128B) and leave as ONE box —
// half as many box rows for the TMA unit, which serves the operand fills of the same CTA (profiles/README.md r6m: the
// TMA stores, not the TMEM reads, the math or the staging, are what the epilogue costs the main loop).
// DUAL (BOXC = 32, 4 KB slab): the pre-activation goes out through `tm_c2` from the slab's first half and the activated
// value through `tm_c` from its second half (GemmParams::C2).  MULPRE: the accumulator is multiplied by gelu'(mul_pre).
// Both are compile-time so that the inference instantiations keep their register budget (102 at 640 threads).
template <int BOXC = 32, bool DUAL = false, bool MULPRE = false, typename F>
__device__ __forceinline__ void gemm_epilogue_tma_bf16_64(const GemmParams& p, const CUtensorMap* tm_c,
                                                          const CUtensorMap* tm_c2, uint32_t taddr, uint32_t slab,
                                                          int row0, int n0, int lane, F after_tmem_reads) {
    static_assert(!DUAL || BOXC == 32, "dual output stages two 32-column boxes");
    if (n0 >= p.N) {
        after_tmem_reads();
        return;
    }
    const uint32_t wrow = slab + lane * (2 * BOXC);
    const int wsw = BOXC == 64 ? (lane & 7) : ((lane >> 1) & 3);
    const bool row_ok = row0 + lane < p.M;
    // LayerNorm fold, consumer side: this lane's row statistics (mu, rstd), finalized by istvt_ln_stats_finalize
    float ln_a = 1.f, ln_b = 0.f;             // C = ln_a * acc + ln_b * c[n] + d[n],  ln_b = -rstd * mu
    if (p.ln_stats != nullptr && row_ok) {
        const float2 st = __ldg(p.ln_stats + row0 + lane);
        ln_a = st.y;
        ln_b = -st.y * st.x;
    }
    float st_n = 0.f, st_sum = 0.f, st_m2 = 0.f, st_shift = 0.f;   // producer side: shifted sums of this warp's columns
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int nb = n0 + hh * 32;
        if (nb >= p.N) {   // warp-uniform: ragged last N tile
            if (hh == 1) after_tmem_reads();
            break;
        }
        uint32_t r[32];
#ifdef ISTVT_GEMM_TRACE
        // timing experiments (results are garbage): trace_no_epi = 2 TMEM reads only; 3 reads + math + staging, no TMA
        // store; 4 staging + TMA store without the TMEM reads; 5 reads + math, nothing staged or stored
        const int tmode = p.trace_no_epi;
        if (tmode == 4) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = static_cast<uint32_t>(lane + j);
        } else
#endif
        tmem_ld_32x32b_x32(taddr + hh * 32, r);
        // the bias of these 32 columns is fetched while the TMEM load is in flight (the same 128 B for every lane)
        const float* bsrc = p.ln_stats != nullptr ? p.ln_d : p.bias;
        float4 bv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
            bv[q] = (bsrc != nullptr && nb + 4 * q < p.N) ? __ldg(reinterpret_cast<const float4*>(bsrc + nb) + q)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        // ff2 data gradient: the pre-activations under this lane's 32 accumulator columns (64 contiguous bytes)
        [[maybe_unused]] uint4 pre[MULPRE ? 4 : 1];
        if constexpr (MULPRE) {
            const uint4* src = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.mul_pre) +
                                                              static_cast<int64_t>(row0 + lane) * p.ld_pre + nb);
#pragma unroll
            for (int g = 0; g < 4; ++g)
                pre[g] = (row_ok && nb + 8 * g < p.N) ? __ldg(src + g) : make_uint4(0u, 0u, 0u, 0u);
        }
        tmem_ld_wait();
        if (hh == 1) after_tmem_reads();
#ifdef ISTVT_GEMM_TRACE
        if (tmode == 2) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) x ^= r[j];
            if (x == 0x7fc12345u) sts_u4(wrow, x, x, x, x);     // keeps the loads alive
            continue;
        }
#endif
        uint32_t o[16];
        [[maybe_unused]] uint32_t o_pre[DUAL ? 16 : 1];
#pragma unroll
        for (int g = 0; g < 4; ++g) {     // 8 columns -> one 16-byte chunk
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
            if constexpr (MULPRE) {
                const uint32_t w[4] = {pre[g].x, pre[g].y, pre[g].z, pre[g].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[2 * j] *= gelu_grad_fast(__uint_as_float(w[j] << 16));
                    v[2 * j + 1] *= gelu_grad_fast(__uint_as_float(w[j] & 0xffff0000u));
                }
            }
            if (p.ln_stats != nullptr) {
                float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
                if (nb + 8 * g < p.N) {
                    c0 = __ldg(reinterpret_cast<const float4*>(p.ln_c + nb) + 2 * g);
                    c1 = __ldg(reinterpret_cast<const float4*>(p.ln_c + nb) + 2 * g + 1);
                }
                v[0] = fmaf(ln_a, v[0], fmaf(ln_b, c0.x, bv[2 * g].x)); v[1] = fmaf(ln_a, v[1], fmaf(ln_b, c0.y, bv[2 * g].y));
                v[2] = fmaf(ln_a, v[2], fmaf(ln_b, c0.z, bv[2 * g].z)); v[3] = fmaf(ln_a, v[3], fmaf(ln_b, c0.w, bv[2 * g].w));
                v[4] = fmaf(ln_a, v[4], fmaf(ln_b, c1.x, bv[2 * g + 1].x)); v[5] = fmaf(ln_a, v[5], fmaf(ln_b, c1.y, bv[2 * g + 1].y));
                v[6] = fmaf(ln_a, v[6], fmaf(ln_b, c1.z, bv[2 * g + 1].z)); v[7] = fmaf(ln_a, v[7], fmaf(ln_b, c1.w, bv[2 * g + 1].w));
            } else {
                v[0] += bv[2 * g].x; v[1] += bv[2 * g].y; v[2] += bv[2 * g].z; v[3] += bv[2 * g].w;
                v[4] += bv[2 * g + 1].x; v[5] += bv[2 * g + 1].y; v[6] += bv[2 * g + 1].z; v[7] += bv[2 * g + 1].w;
            }
            if constexpr (DUAL) {
                o_pre[4 * g] = pack_bf16x2(v[0], v[1]); o_pre[4 * g + 1] = pack_bf16x2(v[2], v[3]);
                o_pre[4 * g + 2] = pack_bf16x2(v[4], v[5]); o_pre[4 * g + 3] = pack_bf16x2(v[6], v[7]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = epi_act(v[j], p.act);
            o[4 * g] = pack_bf16x2(v[0], v[1]); o[4 * g + 1] = pack_bf16x2(v[2], v[3]);
            o[4 * g + 2] = pack_bf16x2(v[4], v[5]); o[4 * g + 3] = pack_bf16x2(v[6], v[7]);
            if (p.row_stats_out != nullptr && nb + 8 * g < p.N) {
                // LayerNorm statistics of the output row (fp32 values: rounding to bf16 moves mean and variance of 728
                // features by ~1e-5 relative).  Sums are taken about a per-row shift — the row's first value in this
                // column group — so that sum of squares does not cancel when |mean| >> sigma.
                if (hh == 0 && g == 0) st_shift = v[0];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float dv = v[j] - st_shift;
                    st_sum += dv;
                    st_m2 = fmaf(dv, dv, st_m2);
                }
                st_n += 8.f;
            }
        }
#ifdef ISTVT_GEMM_TRACE
        if (tmode == 5) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) x ^= o[j];
            if (x == 0x7fc12345u) sts_u4(wrow, x, x, x, x);
            continue;
        }
#endif
        if (BOXC == 32 || hh == 0) {
            if (lane == 0) tma_store_wait_read0();     // the previous box has left the slab
            __syncwarp();
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
            sts_u4(wrow + (DUAL ? 2048 : 0) + (((BOXC == 64 ? hh * 4 + g : g) ^ wsw) << 4), o[4 * g], o[4 * g + 1],
                   o[4 * g + 2], o[4 * g + 3]);
        if constexpr (DUAL) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
                sts_u4(wrow + ((g ^ wsw) << 4), o_pre[4 * g], o_pre[4 * g + 1], o_pre[4 * g + 2], o_pre[4 * g + 3]);
        }
        if (BOXC == 64 && hh == 0 && nb + 32 < p.N) continue;      // the box leaves after the second half
        fence_proxy_async_smem();
        __syncwarp();
#ifdef ISTVT_GEMM_TRACE
        if (tmode == 3) continue;
#endif
        if (lane == 0) {
            if constexpr (DUAL) tma_store_2d(tm_c2, slab, nb, row0);
            tma_store_2d(tm_c, slab + (DUAL ? 2048 : 0), BOXC == 64 ? n0 : nb, row0);
            tma_store_commit();
        }
    }
    if (p.row_stats_out != nullptr && row_ok) {
        // (sum, M2 about the group mean) of this 64-column group: M2 = sum (x - s)^2 - n (mean - s)^2
        const float dmean = st_sum / st_n;
        p.row_stats_out[static_cast<int64_t>(row0 + lane) * ((p.N + 63) >> 6) + (n0 >> 6)] =
            make_float2(fmaf(st_n, st_shift, st_sum), fmaxf(st_m2 - st_sum * dmean, 0.f));
    }
}

// In-place fp32 residual update  C[m, n] += act(acc + bias)  (s_out / ff2 of the inference path, where the GEMM output
// IS the residual stream): the warp stages 32 rows x 32 fp32 columns in its slab in the TMA SWIZZLE_128B layout and one
// elected lane issues a TMA reduce-add of the box — the addition is done by the L2, so the residual never travels to
// the SM.  The register-path epilogue above was the bottleneck of s_out (profiles/README.md r3i: 8 batches of 4
// dependent 16-byte loads per lane and tile, 2 KB in flight per warp, 3.8 TB/s).  Rows >= M and columns >= N are
// clipped by the tensor map.  `row0` = first of the warp's 32 consecutive output rows.
template <typename F>
__device__ __forceinline__ void gemm_epilogue_reduce_64(const GemmParams& p, const CUtensorMap* tm_c, uint32_t taddr,
                                                        uint32_t slab, int row0, int n0, int lane, F after_tmem_reads) {
    if (n0 >= p.N) {
        after_tmem_reads();
        return;
    }
    const uint32_t wrow = slab + lane * 128;
    const int sw = lane & 7;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int nb = n0 + hh * 32;
        if (nb >= p.N) {   // warp-uniform
            if (hh == 1) after_tmem_reads();
            break;
        }
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + hh * 32, r);
        tmem_ld_wait();
        if (hh == 1) after_tmem_reads();
        if (lane == 0) tma_store_wait_read0();     // the previous box has left the slab
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(r[4 * c + j]);
            const int n = nb + c * 4;
            if (p.bias != nullptr && n < p.N) {
                const float4 b = ldg_nc_f4_ordered(p.bias + n);
                v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = epi_act(v[j], p.act);
            sts_u4(wrow + ((c ^ sw) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                   __float_as_uint(v[3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_reduce_add_2d(tm_c, slab, nb, row0);
            tma_store_commit();
        }
    }
}

// host: launch the CTA-pair kernel (gemm_tcgen05_2cta.cu) for a plain GEMM with N >= 256
int launch_gemm_2cta(const void* a, int64_t lda, const void* w, int64_t ldw, const GemmParams& p, cudaStream_t stream);
// MN-major operands (p.mn_major = 1): a = [K rows, M] pitch lda, w = [K rows, N] pitch ldw
int launch_gemm_2cta_mn(const void* a, int64_t lda, const void* w, int64_t ldw, const GemmParams& p, cudaStream_t stream);

}  // namespace istvt
