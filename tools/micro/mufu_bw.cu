// Micro-benchmark: MUFU.EX2 throughput per SM for 1..4 warps per scheduler, (a) bare MUFU chains, (b) the softmax instruction
// mix per pair of scores (FFMA2, 2 x MUFU.EX2, FADD2, 2 x IADD, PRMT) on register-resident data.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I2023-tifs-istvt_b200/csrc tools/micro/mufu_bw.cu -o tools/micro/_bin/mufu_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace istvt;

template <int MODE>
__global__ void __launch_bounds__(512, 1) mufu_kernel(int iters, float seed, long long* clk_out, float* sink) {
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = seed * (i + 1) + threadIdx.x * 1e-6f;
    f32x2_t acc[2] = {0ull, 0ull};
    uint32_t pk = 0;
    const f32x2_t sc2 = f32x2_make(0.999f, 0.999f), nm2 = f32x2_make(-0.001f, -0.001f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = ex2_approx(x[i]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const f32x2_t v = f32x2_fma(f32x2_make(x[2 * j], x[2 * j + 1]), sc2, nm2);
                float a, b;
                f32x2_split(v, a, b);
                const float e0 = ex2_approx(a), e1 = ex2_approx(b);
                acc[j & 1] = f32x2_add(acc[j & 1], f32x2_make(e0, e1));
                pk ^= pack_bf16x2_rne_alu(e0, e1);
                x[2 * j] = e0 - 1.0f; x[2 * j + 1] = e1 - 1.0f;   // keep the chain data dependent, values bounded
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += x[i];
    float a0, a1, a2, a3;
    f32x2_split(acc[0], a0, a1); f32x2_split(acc[1], a2, a3);
    if (s + a0 + a1 + a2 + a3 == 123.456f || pk == 0x1234567u) sink[0] = s;
}

int main() {
    long long* d_clk; float* d_sink;
    cudaMalloc(&d_clk, 148 * sizeof(long long));
    cudaMalloc(&d_sink, 4);
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {4, 8, 12, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) mufu_kernel<0><<<148, warps * 32>>>(iters, 0.01f, d_clk, d_sink);
                else mufu_kernel<1><<<148, warps * 32>>>(iters, 0.01f, d_clk, d_sink);
                cudaDeviceSynchronize();
            }
            long long h[148];
            cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
            const double n = double(warps) * 32 * iters * 32;
            printf("%s warps/SM=%2d (%d per scheduler)  clk=%lld  ex2/clk/SM=%.2f  (%s)\n", mode ? "softmax mix" : "bare MUFU  ", warps,
                   warps / 4, h[0], n / double(h[0]), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
