// HBM-bound row kernels of the transformer: LayerNorm, LayerNorm + temporal self-subtract,
// class-token fill, classification head.  One warp owns one token row (dim <= 1024 kept in registers,
// 16-byte vector accesses, statistics in fp32, two-pass variance like torch).
#include "common.cuh"
#include "simt_util.cuh"

namespace istvt {

constexpr int LN_MAX_ITERS = 8;  // dim <= 8 * 128

// Row held as ITERS x 4 floats per lane: element index = (lane + 32*j)*4 + e
template <int ITERS, typename TI>
__device__ __forceinline__ void ln_load_row(const TI* row, int dim, int lane, float (&v)[ITERS][4]) {
#pragma unroll
    for (int j = 0; j < ITERS; ++j) {
        const int i = (lane + 32 * j) * 4;
        if (i < dim) {
            load4(row + i, v[j]);
        } else {
            v[j][0] = v[j][1] = v[j][2] = v[j][3] = 0.0f;
        }
    }
}

template <int ITERS>
__device__ __forceinline__ void ln_normalize(float (&v)[ITERS][4], int dim, int lane, const float* gamma,
                                             const float* beta, float eps) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < ITERS; ++j) s += (v[j][0] + v[j][1]) + (v[j][2] + v[j][3]);
    const float mean = warp_sum(s) / static_cast<float>(dim);
    float q = 0.0f;
#pragma unroll
    for (int j = 0; j < ITERS; ++j) {
        const int i = (lane + 32 * j) * 4;
        if (i < dim) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = v[j][e] - mean;
                q += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(dim) + eps);
#pragma unroll
    for (int j = 0; j < ITERS; ++j) {
        const int i = (lane + 32 * j) * 4;
        if (i < dim) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + i));
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta + i));
            v[j][0] = (v[j][0] - mean) * rstd * g.x + b.x;
            v[j][1] = (v[j][1] - mean) * rstd * g.y + b.y;
            v[j][2] = (v[j][2] - mean) * rstd * g.z + b.z;
            v[j][3] = (v[j][3] - mean) * rstd * g.w + b.w;
        }
    }
}

template <int ITERS, typename TO>
__device__ __forceinline__ void ln_store_row(TO* row, int dim, int lane, const float (&v)[ITERS][4]) {
#pragma unroll
    for (int j = 0; j < ITERS; ++j) {
        const int i = (lane + 32 * j) * 4;
        if (i < dim) store4(row + i, v[j]);
    }
}

// ------------------------------------------------------------------------------------------
template <int ITERS, typename TI, typename TO>
__global__ void __launch_bounds__(256) layernorm_kernel(const TI* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, TO* __restrict__ y,
                                                        int64_t rows, int dim, int64_t ldx, int64_t ldy, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float v[ITERS][4];
    ln_load_row<ITERS>(x + row * ldx, dim, lane, v);
    ln_normalize<ITERS>(v, dim, lane, gamma, beta, eps);
    ln_store_row<ITERS>(y + row * ldy, dim, lane, v);
}

// One warp per (clip, position): walks the frames, keeps the previous frame's normalised row in
// registers, emits xn and the self-subtract difference (network/vivit/module.py:191-194).
template <int ITERS, typename TO>
__global__ void __launch_bounds__(256)
layernorm_diff_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      TO* __restrict__ xn, TO* __restrict__ diff, int batch, int frames, int tokens, int dim,
                      int64_t ldx, int64_t ldo, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= static_cast<int64_t>(batch) * tokens) return;
    const int b = static_cast<int>(wid / tokens);
    const int pos = static_cast<int>(wid - static_cast<int64_t>(b) * tokens);
    const int64_t row0 = static_cast<int64_t>(b) * frames * tokens + pos;      // token row of frame 0; + f * tokens

    float prev[ITERS][4], cur[ITERS][4], nxt[ITERS][4];
    ln_load_row<ITERS>(x + row0 * ldx, dim, lane, nxt);
    for (int f = 0; f < frames; ++f) {
#pragma unroll
        for (int j = 0; j < ITERS; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) cur[j][e] = nxt[j][e];
        if (f + 1 < frames)
            ln_load_row<ITERS>(x + (row0 + static_cast<int64_t>(f + 1) * tokens) * ldx, dim, lane, nxt);  // prefetch
        ln_normalize<ITERS>(cur, dim, lane, gamma, beta, eps);
        const int64_t off = (row0 + static_cast<int64_t>(f) * tokens) * ldo;
        ln_store_row<ITERS>(xn + off, dim, lane, cur);
        if (f < 2) {
            ln_store_row<ITERS>(diff + off, dim, lane, cur);
        } else {
            float d[ITERS][4];
#pragma unroll
            for (int j = 0; j < ITERS; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) d[j][e] = cur[j][e] - prev[j][e];
            ln_store_row<ITERS>(diff + off, dim, lane, d);
        }
#pragma unroll
        for (int j = 0; j < ITERS; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) prev[j][e] = cur[j][e];
    }
}

// tokens[b, 0, p, :] = temporal_token ; tokens[b, f+1, 0, :] = space_token + pos_emb[f, 0, :]
__global__ void __launch_bounds__(256)
token_fill_kernel(float* __restrict__ tokens, const float* __restrict__ space_token,
                  const float* __restrict__ temporal_token, const float* __restrict__ pos_emb, int batch, int t,
                  int tpf, int dim) {
    const int dim4 = dim >> 2;
    const int64_t per_clip = static_cast<int64_t>(tpf + t) * dim4;  // frame-0 rows + one cls row per real frame
    const int64_t total = per_clip * batch;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / per_clip);
        const int64_t r = i - b * per_clip;
        const int row = static_cast<int>(r / dim4);
        const int c = static_cast<int>(r - static_cast<int64_t>(row) * dim4) * 4;
        const int64_t clip_base = static_cast<int64_t>(b) * (t + 1) * tpf * dim;
        if (row < tpf) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(temporal_token + c));
            *reinterpret_cast<float4*>(tokens + clip_base + static_cast<int64_t>(row) * dim + c) = v;
        } else {
            const int f = row - tpf;
            const float4 s = __ldg(reinterpret_cast<const float4*>(space_token + c));
            const float4 e = __ldg(reinterpret_cast<const float4*>(pos_emb + static_cast<int64_t>(f) * tpf * dim + c));
            *reinterpret_cast<float4*>(tokens + clip_base + static_cast<int64_t>(f + 1) * tpf * dim + c) =
                make_float4(s.x + e.x, s.y + e.y, s.z + e.z, s.w + e.w);
        }
    }
}

// One warp per clip: LN(norm) -> LN(head) -> dot(head_w) + bias, on token row (0, 0).
template <int ITERS>
__global__ void __launch_bounds__(32)
head_kernel(const float* __restrict__ tokens, int64_t rows_per_clip, const float* __restrict__ norm_g,
            const float* __restrict__ norm_b, const float* __restrict__ head_g, const float* __restrict__ head_b,
            const float* __restrict__ head_w, const float* __restrict__ head_bias, float* __restrict__ logits,
            int dim, float eps) {
    const int lane = threadIdx.x;
    const int b = blockIdx.x;
    float v[ITERS][4];
    ln_load_row<ITERS>(tokens + static_cast<int64_t>(b) * rows_per_clip * dim, dim, lane, v);
    ln_normalize<ITERS>(v, dim, lane, norm_g, norm_b, eps);
    ln_normalize<ITERS>(v, dim, lane, head_g, head_b, eps);
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < ITERS; ++j) {
        const int i = (lane + 32 * j) * 4;
        if (i < dim) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(head_w + i));
            acc += v[j][0] * w.x + v[j][1] * w.y + v[j][2] * w.z + v[j][3] * w.w;
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) logits[b] = acc + head_bias[0];
}

template <int ITERS>
static int launch_ln(const void* x, int x_dtype, const float* g, const float* b, void* y, int y_dtype, int64_t rows,
                     int dim, int64_t ldx, int64_t ldy, float eps, cudaStream_t st) {
    const int warps = 8;
    const unsigned grid = static_cast<unsigned>((rows + warps - 1) / warps);
    if (x_dtype == ISTVT_F32 && y_dtype == ISTVT_BF16)
        layernorm_kernel<ITERS, float, __nv_bfloat16><<<grid, warps * 32, 0, st>>>(
            static_cast<const float*>(x), g, b, static_cast<__nv_bfloat16*>(y), rows, dim, ldx, ldy, eps);
    else if (x_dtype == ISTVT_F32 && y_dtype == ISTVT_F32)
        layernorm_kernel<ITERS, float, float><<<grid, warps * 32, 0, st>>>(static_cast<const float*>(x), g, b,
                                                                          static_cast<float*>(y), rows, dim, ldx, ldy, eps);
    else if (x_dtype == ISTVT_BF16 && y_dtype == ISTVT_BF16)
        layernorm_kernel<ITERS, __nv_bfloat16, __nv_bfloat16><<<grid, warps * 32, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), g, b, static_cast<__nv_bfloat16*>(y), rows, dim, ldx, ldy, eps);
    else
        layernorm_kernel<ITERS, __nv_bfloat16, float><<<grid, warps * 32, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), g, b, static_cast<float*>(y), rows, dim, ldx, ldy, eps);
    count_launch();
    return launch_status();
}

template <int ITERS>
static int launch_ln_diff(const float* x, const float* g, const float* b, void* xn, void* diff, int out_dtype,
                          int batch, int frames, int tokens, int dim, int64_t ldx, int64_t ldo, float eps,
                          cudaStream_t st) {
    const int warps = 8;
    const int64_t work = static_cast<int64_t>(batch) * tokens;
    const unsigned grid = static_cast<unsigned>((work + warps - 1) / warps);
    if (out_dtype == ISTVT_BF16)
        layernorm_diff_kernel<ITERS, __nv_bfloat16><<<grid, warps * 32, 0, st>>>(
            x, g, b, static_cast<__nv_bfloat16*>(xn), static_cast<__nv_bfloat16*>(diff), batch, frames, tokens, dim,
            ldx, ldo, eps);
    else
        layernorm_diff_kernel<ITERS, float><<<grid, warps * 32, 0, st>>>(
            x, g, b, static_cast<float*>(xn), static_cast<float*>(diff), batch, frames, tokens, dim, ldx, ldo, eps);
    count_launch();
    return launch_status();
}

#define ISTVT_DISPATCH_ITERS(iters, CALL)  \
    switch (iters) {                       \
        case 1: return CALL(1);            \
        case 2: return CALL(2);            \
        case 3: return CALL(3);            \
        case 4: return CALL(4);            \
        case 5: return CALL(5);            \
        case 6: return CALL(6);            \
        case 7: return CALL(7);            \
        default: return CALL(8);           \
    }

}  // namespace istvt

using namespace istvt;

extern "C" int istvt_layernorm_fwd_ld(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta,
                                      void* y, int y_dtype, int64_t ldy, int64_t rows, int dim, float eps,
                                      istvt_stream_t stream) {
    ISTVT_REQUIRE(x && y && gamma && beta);
    ISTVT_REQUIRE(rows > 0 && dim > 0 && dim % 4 == 0 && dim <= LN_MAX_ITERS * 128);
    ISTVT_REQUIRE((x_dtype | 1) == 1 && (y_dtype | 1) == 1);
    ISTVT_REQUIRE(ldx >= dim && ldy >= dim && ldx % 4 == 0 && ldy % 4 == 0);      // 8-byte (bf16) / 16-byte vectors
    const int iters = (dim + 127) / 128;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(I) launch_ln<I>(x, x_dtype, gamma, beta, y, y_dtype, rows, dim, ldx, ldy, eps, st)
    ISTVT_DISPATCH_ITERS(iters, CALL)
#undef CALL
}

extern "C" int istvt_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y,
                                   int y_dtype, int64_t rows, int dim, float eps, istvt_stream_t stream) {
    return istvt_layernorm_fwd_ld(x, x_dtype, dim, gamma, beta, y, y_dtype, dim, rows, dim, eps, stream);
}

extern "C" int istvt_layernorm_diff_fwd_ld(const float* x, int64_t ldx, const float* gamma, const float* beta, void* xn,
                                           void* diff, int out_dtype, int64_t ld_out, int batch, int frames, int tokens,
                                           int dim, float eps, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && xn && diff && gamma && beta);
    ISTVT_REQUIRE(batch > 0 && frames > 0 && tokens > 0 && dim > 0 && dim % 4 == 0 && dim <= LN_MAX_ITERS * 128);
    ISTVT_REQUIRE((out_dtype | 1) == 1);
    ISTVT_REQUIRE(ldx >= dim && ld_out >= dim && ldx % 4 == 0 && ld_out % 4 == 0);
    const int iters = (dim + 127) / 128;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(I) launch_ln_diff<I>(x, gamma, beta, xn, diff, out_dtype, batch, frames, tokens, dim, ldx, ld_out, eps, st)
    ISTVT_DISPATCH_ITERS(iters, CALL)
#undef CALL
}

extern "C" int istvt_layernorm_diff_fwd(const float* x, const float* gamma, const float* beta, void* xn, void* diff,
                                        int out_dtype, int batch, int frames, int tokens, int dim, float eps,
                                        istvt_stream_t stream) {
    return istvt_layernorm_diff_fwd_ld(x, dim, gamma, beta, xn, diff, out_dtype, dim, batch, frames, tokens, dim, eps,
                                       stream);
}

extern "C" int istvt_token_fill_fwd(float* tokens, const float* space_token, const float* temporal_token,
                                    const float* pos_emb, int batch, int t, int tokens_per_frame, int dim,
                                    istvt_stream_t stream) {
    ISTVT_REQUIRE(tokens && space_token && temporal_token && pos_emb);
    ISTVT_REQUIRE(batch > 0 && t > 0 && tokens_per_frame > 0 && dim > 0 && dim % 4 == 0);
    const int64_t total = static_cast<int64_t>(tokens_per_frame + t) * (dim / 4) * batch;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    token_fill_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        tokens, space_token, temporal_token, pos_emb, batch, t, tokens_per_frame, dim);
    count_launch();
    return launch_status();
}

template <int ITERS>
static int launch_head(const float* tokens, int64_t rpc, const float* ng, const float* nb, const float* hg,
                       const float* hb, const float* hw, const float* hbias, float* logits, int batch, int dim,
                       float eps, cudaStream_t st) {
    head_kernel<ITERS><<<batch, 32, 0, st>>>(tokens, rpc, ng, nb, hg, hb, hw, hbias, logits, dim, eps);
    count_launch();
    return launch_status();
}

extern "C" int istvt_head_fwd(const float* tokens, int64_t rows_per_clip, const float* norm_g, const float* norm_b,
                              const float* head_g, const float* head_b, const float* head_w, const float* head_bias,
                              float* logits, int batch, int dim, float eps, istvt_stream_t stream) {
    ISTVT_REQUIRE(tokens && norm_g && norm_b && head_g && head_b && head_w && head_bias && logits);
    ISTVT_REQUIRE(batch > 0 && dim > 0 && dim % 4 == 0 && dim <= LN_MAX_ITERS * 128 && rows_per_clip > 0);
    const int iters = (dim + 127) / 128;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(I) \
    launch_head<I>(tokens, rows_per_clip, norm_g, norm_b, head_g, head_b, head_w, head_bias, logits, batch, dim, eps, st)
    ISTVT_DISPATCH_ITERS(iters, CALL)
#undef CALL
}
