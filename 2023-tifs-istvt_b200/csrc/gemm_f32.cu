// fp32 SIMT GEMM (FFMA) — the 1e-4 validation mode of the path, not the performance mode.
//   C = act(A · Wᵀ + bias) + residual, A [M,K], W [N,K] row-major, 128x128x16 tiles, 8x8 per thread.
// Shares the "conv mode" of the tensor-core kernel: with taps = 9 the A rows of tap (ky,kx) are the rows
// shifted by ky*w_in + kx and the epilogue compacts the padded output grid (see gemm_tcgen05.cu).
// Also hosts istvt_conv3x3_fwd, which dispatches to the tcgen05 kernel (bf16) or to this one (fp32).
#include "common.cuh"
#include "simt_util.cuh"

namespace istvt {

int conv3x3_bf16(const void* x, const void* wt, const float* bias, void* y, int n, int h, int w, int cin, int cout,
                 int act, cudaStream_t stream);

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16;

struct SgemmParams {
    const float* A; int64_t lda; int64_t a_rows;
    const float* W; int64_t ldw;
    float* C; int64_t ldc;
    int64_t M; int N; int K;
    int taps; int conv_w_in; int conv_h_in;
    const float* bias; const float* residual; int64_t ldr;
    int act;
};

__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmParams p) {
    __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
    __shared__ __align__(16) float Ws[SG_BK][SG_BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = static_cast<int64_t>(blockIdx.x) * SG_BM;
    const int n0 = blockIdx.y * SG_BN;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8 x 8 outputs (strided by 16... no: contiguous 4+4)

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    // loader mapping: 128 rows x 4 float4 per tile = 512 float4, 2 per thread
    const int lrow = tid >> 2;         // 0..63 (+64)
    const int lk = (tid & 3) * 4;      // 0,4,8,12

    for (int tap = 0; tap < p.taps; ++tap) {
        const int64_t shift = p.taps == 1 ? 0 : static_cast<int64_t>(tap / 3) * p.conv_w_in + (tap % 3);
        for (int k0 = 0; k0 < p.K; k0 += SG_BK) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = lrow + r * 64;
                const int k = k0 + lk;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                const int64_t am = m0 + row + shift;
                if (am < p.a_rows && k < p.K) a = *reinterpret_cast<const float4*>(p.A + am * p.lda + k);
                const int wn = n0 + row;
                if (wn < p.N && k < p.K)
                    b = *reinterpret_cast<const float4*>(p.W + static_cast<int64_t>(wn) * p.ldw + tap * p.K + k);
                As[lk][row] = a.x; As[lk + 1][row] = a.y; As[lk + 2][row] = a.z; As[lk + 3][row] = a.w;
                Ws[lk][row] = b.x; Ws[lk + 1][row] = b.y; Ws[lk + 2][row] = b.z; Ws[lk + 3][row] = b.w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < SG_BK; ++kk) {
                float a[8], b[8];
                const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Ws[kk][64 + tx * 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= p.M) continue;
        int64_t drow = m;
        if (p.taps != 1) {
            const int64_t img_sz = static_cast<int64_t>(p.conv_h_in) * p.conv_w_in;
            const int64_t img = m / img_sz;
            const int rem = static_cast<int>(m - img * img_sz);
            const int y = rem / p.conv_w_in, x = rem - y * p.conv_w_in;
            const int ho = p.conv_h_in - 2, wo = p.conv_w_in - 2;
            if (y >= ho || x >= wo) continue;
            drow = (img * ho + y) * wo + x;
        }
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 64 + tx * 4;
            if (n >= p.N) continue;
            float v[4] = {acc[i][jh * 4], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
            if (p.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
            }
            if (p.act == ISTVT_ACT_RELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.0f);
            } else if (p.act == ISTVT_ACT_GELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = 0.5f * v[e] * (1.0f + erff(v[e] * 0.70710678118654752440f));
            }
            if (p.residual) {
                const float4 r = *reinterpret_cast<const float4*>(p.residual + drow * p.ldr + n);
                v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
            }
            *reinterpret_cast<float4*>(p.C + drow * p.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

static int launch_sgemm(const SgemmParams& p, cudaStream_t st) {
    ISTVT_REQUIRE(p.A && p.W && p.C);
    ISTVT_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0 && p.N % 4 == 0 && p.K % 4 == 0);
    ISTVT_REQUIRE(p.lda % 4 == 0 && p.ldw % 4 == 0 && p.ldc % 4 == 0 && (p.residual == nullptr || p.ldr % 4 == 0));
    const int64_t m_tiles = (p.M + SG_BM - 1) / SG_BM;
    ISTVT_REQUIRE(m_tiles < (int64_t(1) << 31));
    dim3 grid(static_cast<unsigned>(m_tiles), (p.N + SG_BN - 1) / SG_BN);
    sgemm_kernel<<<grid, 256, 0, st>>>(p);
    count_launch();
    return launch_status();
}

}  // namespace istvt

using namespace istvt;

extern "C" int istvt_gemm_f32_fwd(const float* a, int64_t lda, const float* w, int64_t ldw, float* c, int64_t ldc,
                                  int64_t m, int n, int k, const float* bias, const float* residual, int64_t ldr,
                                  int act, istvt_stream_t stream) {
    ISTVT_REQUIRE(act >= ISTVT_ACT_NONE && act <= ISTVT_ACT_GELU);
    SgemmParams p{};
    p.A = a; p.lda = lda; p.a_rows = m;
    p.W = w; p.ldw = ldw;
    p.C = c; p.ldc = ldc;
    p.M = m; p.N = n; p.K = k;
    p.taps = 1; p.conv_w_in = 0; p.conv_h_in = 0;
    p.bias = bias; p.residual = residual; p.ldr = ldr;
    p.act = act;
    return launch_sgemm(p, static_cast<cudaStream_t>(stream));
}

extern "C" int istvt_conv3x3_fwd(const void* x, const void* wt, const float* bias, void* y, int dtype, int n, int h,
                                 int w, int cin, int cout, int act, istvt_stream_t stream) {
    ISTVT_REQUIRE(x && wt && y);
    ISTVT_REQUIRE(n > 0 && h >= 3 && w >= 3 && cin > 0 && cout > 0);
    ISTVT_REQUIRE(act >= ISTVT_ACT_NONE && act <= ISTVT_ACT_GELU);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == ISTVT_BF16) return conv3x3_bf16(x, wt, bias, y, n, h, w, cin, cout, act, st);
    ISTVT_REQUIRE(dtype == ISTVT_F32);
    SgemmParams p{};
    p.A = static_cast<const float*>(x); p.lda = cin; p.a_rows = static_cast<int64_t>(n) * h * w;
    p.W = static_cast<const float*>(wt); p.ldw = 9 * static_cast<int64_t>(cin);
    p.C = static_cast<float*>(y); p.ldc = cout;
    p.M = static_cast<int64_t>(n) * h * w; p.N = cout; p.K = cin;
    p.taps = 9; p.conv_w_in = w; p.conv_h_in = h;
    p.bias = bias; p.residual = nullptr; p.ldr = 0;
    p.act = act;
    return launch_sgemm(p, st);
}
