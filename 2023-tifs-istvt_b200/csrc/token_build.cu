// Token-sequence assembly of the ablation transformers: class token in slot 0,
// patch rows behind it, positional embedding added — the `repeat` / `torch.cat` / `x += pos_embedding` triples of
//   ViViT.forward      vivit.py:64-66  (sequence = one frame: space token + 361 patches, pos_embedding[f, slot])
//   ViViT.forward      vivit.py:73-74  (sequence = one clip: temporal token + T frame tokens, no positional term)
//   VanillaTr.forward  vivit.py:183-185 (sequence = one clip: class token + T*361 patches, pos_embedding[slot])
// as ONE pass: read the patch rows (bf16 or fp32), write the fp32 residual stream.  HBM-bound: in + out bytes.
#include "common.cuh"
#include "simt_util.cuh"

namespace istvt {

// tokens[seq, 0, :]     = cls + pos[(seq % pos_period), 0, :]
// tokens[seq, 1 + i, :] = src[seq * n + i, :] + pos[(seq % pos_period), 1 + i, :]
// one thread per 4 channels of one output row.
template <typename T>
__global__ void __launch_bounds__(256)
token_build_kernel(const T* __restrict__ src, const float* __restrict__ cls, const float* __restrict__ pos,
                   float* __restrict__ tokens, int64_t total4, int n, int dim4, int pos_period) {
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total4;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int c4 = static_cast<int>(idx % dim4);
        const int64_t row = idx / dim4;
        const int slot = static_cast<int>(row % (n + 1));
        const int64_t seq = row / (n + 1);
        float v[4];
        if (slot == 0) {
            load4(cls + c4 * 4, v);
        } else {
            load4(src + ((seq * n + slot - 1) * dim4 + c4) * 4, v);
        }
        if (pos != nullptr) {
            float p[4];
            load4(pos + ((static_cast<int64_t>(seq % pos_period) * (n + 1) + slot) * dim4 + c4) * 4, p);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] += p[e];
        }
        store4(tokens + idx * 4, v);
    }
}

// out[seq, :] = mean over the n rows of x[seq, :, :] (fp32): ViViT's `x.mean(dim = 1)` pooling, vivit.py:79.
// One thread per 4 channels of one sequence; n is small (T + 1 frame tokens).
__global__ void __launch_bounds__(256)
mean_rows_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t total4, int n, int dim4) {
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total4;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int c4 = static_cast<int>(idx % dim4);
        const int64_t seq = idx / dim4;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int r = 0; r < n; ++r) {
            float v[4];
            load4(x + ((seq * n + r) * dim4 + c4) * 4, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] += v[e];
        }
        const float inv = 1.0f / static_cast<float>(n);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] *= inv;
        store4(out + idx * 4, acc);
    }
}

}  // namespace istvt

using namespace istvt;

extern "C" int istvt_mean_rows_fwd(const float* x, float* out, int sequences, int n, int dim, istvt_stream_t stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ISTVT_REQUIRE(x && out);
    ISTVT_REQUIRE(sequences > 0 && n > 0 && dim > 0 && dim % 4 == 0);
    const int dim4 = dim / 4;
    const int64_t total4 = static_cast<int64_t>(sequences) * dim4;
    int64_t blocks = (total4 + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    mean_rows_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(x, out, total4, n, dim4);
    count_launch();
    return launch_status();
}

extern "C" int istvt_token_build_fwd(const void* src, int src_dtype, const float* cls, const float* pos, float* tokens,
                                     int sequences, int n, int dim, int pos_period, istvt_stream_t stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ISTVT_REQUIRE(src && cls && tokens);
    ISTVT_REQUIRE(sequences > 0 && n > 0 && dim > 0 && dim % 4 == 0 && pos_period > 0);
    ISTVT_REQUIRE(src_dtype == ISTVT_BF16 || src_dtype == ISTVT_F32);
    const int dim4 = dim / 4;
    const int64_t total4 = static_cast<int64_t>(sequences) * (n + 1) * dim4;
    int64_t blocks = (total4 + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    if (src_dtype == ISTVT_BF16)
        token_build_kernel<__nv_bfloat16><<<static_cast<int>(blocks), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(src), cls, pos, tokens, total4, n, dim4, pos_period);
    else
        token_build_kernel<float><<<static_cast<int>(blocks), 256, 0, st>>>(
            static_cast<const float*>(src), cls, pos, tokens, total4, n, dim4, pos_period);
    count_launch();
    return launch_status();
}
