// Xception stem conv1 (3x3, stride 2, no padding, 3 -> 32 channels) + folded BatchNorm + ReLU as an implicit GEMM on the
// tcgen05 tensor cores in TF32:  network/xception.py:118-120,194-196.
//
// Why: the SIMT kernel (entry_flow.cu, kept for the fp32 validation mode) spends 74 % of its instructions on the
// 27 x 32 FFMAs per output pixel and runs at 0.23 of the HBM roofline (ncu, profiles/r6a_conv_stem_ncu_source.txt:
// issue 54 %, FFMA-bound).  Here the multiply-adds move to the tensor pipe and the SIMT side only gathers:
//   * tile = 128 consecutive output pixels of the flattened (image, row, column) index = the 128 TMEM lanes;
//   * thread t builds row t of the A operand: its pixel's 27 taps as fp32 (= TF32 operand bits; K padded to 32), one
//     128-byte row of a K-major SWIZZLE_128B tile, written with 8 conflict-free 16-byte stores;
//   * the 32 x 32 weight matrix (K-major, SW128, 4 KB) stays in shared memory for the life of the CTA;
//   * one thread issues 4 tcgen05.mma kind::tf32 (M 128, N 32, K 8) into one of two 32-column TMEM accumulators and
//     commits to an mbarrier; the epilogue of tile i (tcgen05.ld: thread t <- lane t = its own pixel, + bias, ReLU,
//     bf16, one contiguous 64-byte NHWC pixel per thread) runs while the MMAs of tile i+1 are in flight;
//   * ~37 KB of shared memory and 64 TMEM columns per CTA: five CTAs per SM hide the gather latency.
// TF32 keeps 10 mantissa bits of inputs and weights (the bf16 output keeps 8); accumulation is fp32.
// The uint8 NHWC variant (decoded frames, normalisation folded into the weights by the host) differs only in the gather
// and in the K order (ky, kx, c instead of c, ky, kx).
#include "common.cuh"
#include "ptx.cuh"

namespace istvt {

constexpr int STC_CO = 32;
constexpr int STC_TILE = 128;
constexpr int STC_A_BYTES = STC_TILE * 128;          // 128 rows x 32 fp32
constexpr int STC_W_BYTES = STC_CO * 128;
constexpr int STC_SMEM = 2 * STC_A_BYTES + STC_W_BYTES + 1024 /*align*/ + 256 /*bias, barriers, TMEM holder*/;

__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// instruction descriptor, kind::tf32: D fp32, A / B TF32 (format 2), both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

template <bool U8>
__global__ void __launch_bounds__(128)
conv_stem_tc_kernel(const void* __restrict__ xin, const float* __restrict__ wt, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ y, int n, int h, int w, int ho, int wo, int relu) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                   // 2 x 16 KB
    uint8_t* smem_w = smem + 2 * STC_A_BYTES;                 // 4 KB
    float* s_bias = reinterpret_cast<float*>(smem_w + STC_W_BYTES);
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(s_bias + STC_CO);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(mma_done + 2);

    const int t = threadIdx.x;
    const int warp = t >> 5;
    if (t == 0) {
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_holder, 64);
        tmem_relinquish();
    }
    // weights -> K-major SW128 rows: row co, k = (c, ky, kx) [fp32 input] or (ky, kx, c) [uint8 input], k 27..31 zero
    for (int i = t; i < STC_CO * 32; i += 128) {
        const int co = i >> 5, k = i & 31;
        float v = 0.f;
        if (k < 27) {
            int src = k;                                       // wt is [co][c][ky][kx]
            if (U8) {
                const int ky = k / 9, kx = (k - ky * 9) / 3, c = k % 3;
                src = c * 9 + ky * 3 + kx;
            }
            v = wt[co * 27 + src];
        }
        *reinterpret_cast<float*>(smem_w + co * 128 + ((((k >> 2) ^ (co & 7)) << 4) | ((k & 3) << 2))) = v;
    }
    if (t < STC_CO) s_bias[t] = bias[t];
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    const int64_t total = static_cast<int64_t>(n) * ho * wo;
    const int64_t tiles = (total + STC_TILE - 1) / STC_TILE;
    const uint64_t desc_hi = make_smem_desc(0, 0, 1024, SWZ_128B);
    const uint32_t a_field0 = (smem_u32(smem_a) & 0x3FFFFu) >> 4;
    const uint64_t b_desc = desc_hi | ((smem_u32(smem_w) & 0x3FFFFu) >> 4);
    constexpr uint32_t idesc = make_idesc_tf32(STC_TILE, STC_CO);
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;

    auto epilogue = [&](int64_t tile, int buf, uint32_t phase) {
        mbar_wait(&mma_done[buf], phase);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + lane_base + buf * STC_CO, r);
        tmem_ld_wait();
        tc_fence_before();
        const int64_t p = tile * STC_TILE + t;
        if (p < total) {
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float v0 = __uint_as_float(r[2 * j]) + s_bias[2 * j];
                float v1 = __uint_as_float(r[2 * j + 1]) + s_bias[2 * j + 1];
                if (relu) {
                    v0 = fmaxf(v0, 0.f);
                    v1 = fmaxf(v1, 0.f);
                }
                o[j] = pack_bf16x2(v0, v1);
            }
            uint4* dst = reinterpret_cast<uint4*>(y + p * STC_CO);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
    };

    int it = 0;
    int64_t prev_tile = -1;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        // ---- gather this thread's pixel: 27 taps -> one 128-byte A row ----
        float a[32];
#pragma unroll
        for (int k = 27; k < 32; ++k) a[k] = 0.f;
        const int64_t p = tile * STC_TILE + t;
        if (p < total) {
            const int ox = static_cast<int>(p % wo);
            const int64_t q = p / wo;
            const int oy = static_cast<int>(q % ho);
            const int64_t img = q / ho;
            if (!U8) {
                const float* x = static_cast<const float*>(xin);
                const bool even = (w & 1) == 0;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const float* src = x + ((img * 3 + c) * h + 2 * oy + ky) * static_cast<int64_t>(w) + 2 * ox;
                        if (even) {     // 2 ox and w even: the row start is 8-byte aligned
                            const float2 v01 = __ldg(reinterpret_cast<const float2*>(src));
                            a[c * 9 + ky * 3] = v01.x;
                            a[c * 9 + ky * 3 + 1] = v01.y;
                        } else {
                            a[c * 9 + ky * 3] = __ldg(src);
                            a[c * 9 + ky * 3 + 1] = __ldg(src + 1);
                        }
                        a[c * 9 + ky * 3 + 2] = __ldg(src + 2);
                    }
            } else {
                const uint8_t* x = static_cast<const uint8_t*>(xin);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const uint8_t* src = x + ((img * h + 2 * oy + ky) * static_cast<int64_t>(w) + 2 * ox) * 3;
#pragma unroll
                    for (int j = 0; j < 9; ++j) a[ky * 9 + j] = static_cast<float>(__ldg(src + j));
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 27; ++k) a[k] = 0.f;
        }
        // A[buf] was last read by the MMAs of iteration it - 2, whose completion every thread observed in its epilogue
        const uint32_t arow = smem_u32(smem_a + buf * STC_A_BYTES) + t * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            sts_u4(arow + ((j ^ (t & 7)) << 4), __float_as_uint(a[4 * j]), __float_as_uint(a[4 * j + 1]),
                   __float_as_uint(a[4 * j + 2]), __float_as_uint(a[4 * j + 3]));
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
            const uint64_t a_desc = desc_hi | (a_field0 + buf * (STC_A_BYTES >> 4));
            const uint32_t d_tmem = tmem_base + buf * STC_CO;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32_ss(d_tmem, a_desc + k * 2, b_desc + k * 2, idesc, k != 0 ? 1u : 0u);
            umma_commit(&mma_done[buf]);
        }
        // ---- epilogue of the previous tile while these MMAs run ----
        if (prev_tile >= 0) epilogue(prev_tile, buf ^ 1, static_cast<uint32_t>(((it - 1) >> 1) & 1));
        prev_tile = tile;
    }
    if (prev_tile >= 0) epilogue(prev_tile, (it - 1) & 1, static_cast<uint32_t>(((it - 1) >> 1) & 1));

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

// host: bf16-output stem on the tensor cores; returns ISTVT_OK or an error code
int conv_stem_tc_launch(const void* x, bool u8, const float* wt, const float* bias, void* y, int n, int h, int w,
                        int relu, cudaStream_t st) {
    const int ho = (h - 3) / 2 + 1, wo = (w - 3) / 2 + 1;
    const int64_t tiles = (static_cast<int64_t>(n) * ho * wo + STC_TILE - 1) / STC_TILE;
    int64_t blocks = static_cast<int64_t>(sm_count()) * 5;
    if (blocks > tiles) blocks = tiles;
    if (u8) {
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(conv_stem_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, STC_SMEM));
        conv_stem_tc_kernel<true><<<static_cast<unsigned>(blocks), 128, STC_SMEM, st>>>(
            x, wt, bias, static_cast<__nv_bfloat16*>(y), n, h, w, ho, wo, relu);
    } else {
        ISTVT_CHECK_CUDA(cudaFuncSetAttribute(conv_stem_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, STC_SMEM));
        conv_stem_tc_kernel<false><<<static_cast<unsigned>(blocks), 128, STC_SMEM, st>>>(
            x, wt, bias, static_cast<__nv_bfloat16*>(y), n, h, w, ho, wo, relu);
    }
    count_launch();
    return launch_status();
}

}  // namespace istvt
