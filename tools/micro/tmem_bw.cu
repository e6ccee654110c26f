// Micro-benchmark: tcgen05.ld throughput per SM (bytes / clk) for 4, 8 and 16 reading warps (x32 and x16 shapes).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I2023-tifs-istvt_b200/csrc tools/micro/tmem_bw.cu -o gpurun_out/tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace istvt;

template <int X>
__global__ void __launch_bounds__(640, 1) tmem_read_kernel(int warps, int iters, long long* clk_out, uint32_t* sink) {
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = holder;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < warps) {
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int part = warp >> 2;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (X == 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(base + lane_base + ((part * 32 + c * 128) & 511), r);
                    tmem_ld_wait();
                    acc ^= r[0] ^ r[31];
                } else {
                    uint32_t r[16];
                    tmem_ld_32x32b_x16(base + lane_base + ((part * 16 + c * 64) & 511), r);
                    tmem_ld_wait();
                    acc ^= r[0] ^ r[15];
                }
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}

// no wait between the four loads of an iteration (4 loads in flight per warp)
__global__ void __launch_bounds__(640, 1) tmem_read4_kernel(int warps, int iters, long long* clk_out, uint32_t* sink) {
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = holder;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < warps) {
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        for (int i = 0; i < iters; ++i) {
            uint32_t r0[32], r1[32], r2[32], r3[32];
            tmem_ld_32x32b_x32(base + lane_base + 0, r0);
            tmem_ld_32x32b_x32(base + lane_base + 32, r1);
            tmem_ld_32x32b_x32(base + lane_base + 64, r2);
            tmem_ld_32x32b_x32(base + lane_base + 96, r3);
            tmem_ld_wait();
            acc ^= r0[0] ^ r1[31] ^ r2[5] ^ r3[7];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}

int main() {
    long long* d_clk; uint32_t* d_sink;
    cudaMalloc(&d_clk, 148 * sizeof(long long));
    cudaMalloc(&d_sink, 4);
    const int iters = 2000;
    for (int x : {32, 16}) {
        for (int warps : {4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (x == 32) tmem_read_kernel<32><<<148, 640>>>(warps, iters, d_clk, d_sink);
                else tmem_read_kernel<16><<<148, 640>>>(warps, iters, d_clk, d_sink);
                cudaDeviceSynchronize();
            }
            long long h[148];
            cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
            const double bytes = double(warps) * iters * 4 * 32 * x * 4;
            printf("x%d warps=%2d  clk=%lld  bytes/clk/SM=%.1f  (%s)\n", x, warps, h[0], bytes / double(h[0]), cudaGetErrorString(cudaGetLastError()));
        }
    }
    for (int warps : {4, 8}) {
        for (int rep = 0; rep < 2; ++rep) { tmem_read4_kernel<<<148, 640>>>(warps, iters, d_clk, d_sink); cudaDeviceSynchronize(); }
        long long h[148];
        cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
        const double bytes = double(warps) * iters * 4 * 32 * 32 * 4;
        printf("4 x x32 in flight, warps=%2d  clk=%lld  bytes/clk/SM=%.1f  (%s)\n", warps, h[0], bytes / double(h[0]), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
