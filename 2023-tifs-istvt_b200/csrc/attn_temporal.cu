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
    ISTVT_REQUIRE(heads * frames <= 288);
    ISTVT_REQUIRE(static_cast<int64_t>(batch) * tokens < (int64_t(1) << 31));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == ISTVT_BF16)
        return launch_temporal<__nv_bfloat16>(qk, v, out, probs, batch, frames, tokens, heads, scale, st);
    if (dtype == ISTVT_F32) return launch_temporal<float>(qk, v, out, probs, batch, frames, tokens, heads, scale, st);
    return ISTVT_ERR_INVALID_ARG;
}
