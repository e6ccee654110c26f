// Xception stem conv2 (3x3, stride 1, no padding, 32 -> 64 channels) + folded BatchNorm + ReLU (xception.py:122-123,
// 198-200) as an implicit GEMM whose A operand is GATHERED by the SIMT threads instead of fetched per tap by TMA.
//
// Why: with 32 input channels every pixel is a 64-byte row, and the 9-tap TMA formulation (gemm_tcgen05.cu, conv mode —
// still the path for every other channel count, e.g. the data-gradient convolution of the training step) asks the TMA
// unit for 9 x 128 box rows of 64 bytes per 128-pixel tile: the kernel ran at the unit's row rate, 0.21 of the HBM
// roofline (profiles/README.md r4d / r4e).  Here, as in conv_stem_tc.cu:
//   * tile = 128 consecutive output pixels of the flattened (image, row, column) index = the 128 TMEM lanes; thread t
//     owns pixel t: per filter row ky it loads the 3 x 64 contiguous bytes under its pixel (12 16-byte loads; neighbours
//     overlap, L1 serves the re-reads) and writes them as row t of three K-major SWIZZLE_64B tap tiles (8 KB each);
//   * the 64 x 288 weight matrix stays in shared memory (9 tap tiles of 4 KB) for the life of the CTA;
//   * three passes (ky = 0, 1, 2) of 6 tcgen05.mma (M 128, N 64, K 16) accumulate one 64-column TMEM tile; the loads
//     of pass k + 1 are issued before the wait for pass k's MMAs, and the previous tile's epilogue (tcgen05.ld: thread
//     t <- its own pixel, + bias, ReLU, bf16, 128 contiguous bytes per thread) runs under the loads of pass 0;
//   * 24 KB (A) + 36 KB (W) of shared memory and 64 TMEM columns per CTA: three CTAs per SM.
#include "common.cuh"
#include "ptx.cuh"

namespace istvt {

constexpr int C2_CI = 32, C2_CO = 64;
constexpr int C2_TILE = 128;
constexpr int C2_A_TAP = C2_TILE * 64;               // 8 KB: 128 pixels x 32 bf16
constexpr int C2_W_TAP = C2_CO * 64;                 // 4 KB
constexpr int C2_SMEM = 3 * C2_A_TAP + 9 * C2_W_TAP + 1024 /*align*/ + 512 /*bias, barrier, TMEM holder*/;

__device__ __forceinline__ uint4 c2_ldg(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(128)
conv3x3_c32_tc_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wt,
                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int n, int h, int w, int act) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                   // 3 tap tiles of the current filter row
    uint8_t* smem_w = smem + 3 * C2_A_TAP;                    // 9 tap tiles
    float* s_bias = reinterpret_cast<float*>(smem_w + 9 * C2_W_TAP);
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(s_bias + C2_CO);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(mma_done + 1);

    const int t = threadIdx.x;
    const int warp = t >> 5;
    if (t == 0) {
        mbar_init(mma_done, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_holder, 64);
        tmem_relinquish();
    }
    // weights [co][ky][kx][ci] -> per tap a K-major SW64 tile: row co, 16-byte chunk q at co*64 + ((q ^ ((co >> 1) & 3)) << 4)
    for (int i = t; i < 9 * C2_CO * 4; i += 128) {
        const int q = i & 3, tap = (i >> 2) % 9, co = i / 36;
        const uint4 v = *reinterpret_cast<const uint4*>(wt + (co * 9 + tap) * C2_CI + q * 8);
        *reinterpret_cast<uint4*>(smem_w + tap * C2_W_TAP + co * 64 + ((q ^ ((co >> 1) & 3)) << 4)) = v;
    }
    if (t < C2_CO) s_bias[t] = bias[t];
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    const int ho = h - 2, wo = w - 2;
    const int64_t total = static_cast<int64_t>(n) * ho * wo;
    const int64_t tiles = (total + C2_TILE - 1) / C2_TILE;
    const uint64_t desc_hi = make_smem_desc(0, 0, 512, SWZ_64B);
    const uint32_t a_field0 = (smem_u32(smem_a) & 0x3FFFFu) >> 4;
    const uint32_t w_field0 = (smem_u32(smem_w) & 0x3FFFFu) >> 4;
    constexpr uint32_t idesc = make_idesc_bf16(C2_TILE, C2_CO, 0, 0);
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t arow = smem_u32(smem_a) + t * 64;
    const int asw = (t >> 1) & 3;

    auto epilogue = [&](int64_t tile) {       // the tile's last commit has been observed by the caller
        tc_fence_after();
        const int64_t p = tile * C2_TILE + t;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + lane_base + hh * 32, r);
            tmem_ld_wait();
            if (p < total) {
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float v0 = __uint_as_float(r[2 * j]) + s_bias[hh * 32 + 2 * j];
                    float v1 = __uint_as_float(r[2 * j + 1]) + s_bias[hh * 32 + 2 * j + 1];
                    if (act == ISTVT_ACT_RELU) {
                        v0 = fmaxf(v0, 0.f);
                        v1 = fmaxf(v1, 0.f);
                    }
                    o[j] = pack_bf16x2(v0, v1);
                }
                uint4* dst = reinterpret_cast<uint4*>(y + p * C2_CO + hh * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            }
        }
        tc_fence_before();
    };

    uint32_t npass = 0;                        // commits issued so far (one per filter row)
    int64_t prev_tile = -1;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t p = tile * C2_TILE + t;
        const bool valid = p < total;
        const __nv_bfloat16* src0 = x;
        if (valid) {
            const int ox = static_cast<int>(p % wo);
            const int64_t q = p / wo;
            const int oy = static_cast<int>(q % ho);
            const int64_t img = q / ho;
            src0 = x + ((img * h + oy) * static_cast<int64_t>(w) + ox) * C2_CI;
        }
#pragma unroll 1
        for (int ky = 0; ky < 3; ++ky) {
            // the 3 x 64 contiguous bytes of filter row ky under this pixel
            uint4 v[12];
            if (valid) {
                const __nv_bfloat16* src = src0 + static_cast<int64_t>(ky) * w * C2_CI;
#pragma unroll
                for (int j = 0; j < 12; ++j) v[j] = c2_ldg(src + j * 8);
            } else {
#pragma unroll
                for (int j = 0; j < 12; ++j) v[j] = make_uint4(0u, 0u, 0u, 0u);
            }
            if (npass > 0) mbar_wait(mma_done, (npass - 1) & 1);     // the previous pass's MMAs have read the A tiles
            if (ky == 0 && prev_tile >= 0) epilogue(prev_tile);      // ... and, if it closed a tile, filled the accumulator
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    sts_u4(arow + kx * C2_A_TAP + ((q ^ asw) << 4), v[4 * kx + q].x, v[4 * kx + q].y, v[4 * kx + q].z,
                           v[4 * kx + q].w);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();
            if (t == 0) {
                tc_fence_after();
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const uint64_t a_desc = desc_hi | (a_field0 + kx * (C2_A_TAP >> 4));
                    const uint64_t b_desc = desc_hi | (w_field0 + (ky * 3 + kx) * (C2_W_TAP >> 4));
                    umma_f16_ss(tmem_base, a_desc, b_desc, idesc, (ky | kx) != 0 ? 1u : 0u);
                    umma_f16_ss(tmem_base, a_desc + 2, b_desc + 2, idesc, 1u);
                }
                umma_commit(mma_done);
            }
            ++npass;
        }
        prev_tile = tile;
    }
    if (prev_tile >= 0) {
        mbar_wait(mma_done, (npass - 1) & 1);
        epilogue(prev_tile);
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

// host: conv2-shaped 3x3 convolution (cin 32, cout 64, bf16) on the gathered-operand kernel
int conv3x3_c32_tc_launch(const void* x, const void* wt, const float* bias, void* y, int n, int h, int w, int act,
                          cudaStream_t st) {
    const int64_t tiles = (static_cast<int64_t>(n) * (h - 2) * (w - 2) + C2_TILE - 1) / C2_TILE;
    int64_t blocks = static_cast<int64_t>(sm_count()) * 3;
    if (blocks > tiles) blocks = tiles;
    ISTVT_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_c32_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM));
    conv3x3_c32_tc_kernel<<<static_cast<unsigned>(blocks), 128, C2_SMEM, st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(wt), bias,
        static_cast<__nv_bfloat16*>(y), n, h, w, act);
    count_launch();
    return launch_status();
}

}  // namespace istvt
