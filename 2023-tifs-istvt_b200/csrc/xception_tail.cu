// Kernels that complete the per-frame Xception baseline (the middle and exit flow
// reuse the entry flow's depthwise / pointwise / pool kernels; only these two ops are new):
//   add            identity-skip residual of the middle-flow blocks 4-11, x += inp           (xception.py:97-101)
//   pool_linear    ReLU + adaptive_avg_pool2d(1,1) + last_linear                             (xception.py:208-221)
#include "common.cuh"
#include "simt_util.cuh"

namespace istvt {

// y = a + b on 8-element vectors; HBM-bound, 3 x elem bytes per element.
template <typename T>
__global__ void __launch_bounds__(256)
add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int64_t n8) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float va[8], vb[8];
        load8(a + i * 8, va);
        load8(b + i * 8, vb);
#pragma unroll
        for (int e = 0; e < 8; ++e) va[e] += vb[e];
        store8(y + i * 8, va);
    }
}

// One CTA per image: mean over the hw pixels of relu(x[img, :, c]) for every channel (threads stride the channels,
// so every global access is a coalesced row segment), then `ncls` dot products reduced across the CTA.
template <typename T>
__global__ void __launch_bounds__(256)
pool_linear_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ out, int hw, int c, int ncls, int relu) {
    extern __shared__ float s_mean[];          // [c]
    __shared__ float s_red[8];
    const T* xin = x + static_cast<int64_t>(blockIdx.x) * hw * c;
    const float inv = 1.0f / static_cast<float>(hw);
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float s = 0.0f;
        for (int p = 0; p < hw; ++p) {
            const float v = to_f32(xin[static_cast<int64_t>(p) * c + ch]);
            s += relu ? fmaxf(v, 0.0f) : v;
        }
        s_mean[ch] = s * inv;
    }
    __syncthreads();
    for (int k = 0; k < ncls; ++k) {
        float part = 0.0f;
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) part = fmaf(s_mean[ch], w[static_cast<int64_t>(k) * c + ch], part);
        part = warp_sum(part);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = bias ? bias[k] : 0.0f;
            for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) tot += s_red[i];
            out[static_cast<int64_t>(blockIdx.x) * ncls + k] = tot;
        }
        __syncthreads();
    }
}

}  // namespace istvt

using namespace istvt;

extern "C" int istvt_add_fwd(const void* a, const void* b, void* y, int dtype, int64_t count, istvt_stream_t stream) {
    ISTVT_REQUIRE(a && b && y && count > 0 && count % 8 == 0);
    ISTVT_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t n8 = count / 8;
    int64_t blocks = (n8 + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == ISTVT_BF16)
        add_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(y), n8);
    else if (dtype == ISTVT_F32)
        add_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const float*>(a),
                                                                          static_cast<const float*>(b), static_cast<float*>(y), n8);
    else
        return ISTVT_ERR_INVALID_ARG;
    count_launch();
    return launch_status();
}

extern "C" int istvt_pool_linear_fwd(const void* x, int dtype, const float* w, const float* bias, float* out, int n,
                                     int hw, int c, int ncls, int relu, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && w && out);
    ISTVT_REQUIRE(n > 0 && hw > 0 && c > 0 && ncls > 0 && c <= 12288);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = static_cast<size_t>(c) * sizeof(float);
    if (dtype == ISTVT_BF16)
        pool_linear_kernel<__nv_bfloat16><<<n, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(x), w, bias, out, hw, c, ncls, relu);
    else if (dtype == ISTVT_F32)
        pool_linear_kernel<float><<<n, 256, smem, st>>>(static_cast<const float*>(x), w, bias, out, hw, c, ncls, relu);
    else
        return ISTVT_ERR_INVALID_ARG;
    count_launch();
    return launch_status();
}
