// Microbenchmark: how fast does ONE SM's TMA unit fill shared memory, as a function of the box row length and of
// whether the box rows are contiguous in global memory?  (Round-3 question: is the 0.47 us per GEMM k-block a
// byte-rate or a per-row cost?  profiles/README.md r4e.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I2023-tifs-istvt_b200/csrc -Iinclude \
//        tools/micro/tma_fill.cu 2023-tifs-istvt_b200/csrc/common.cu -o gpurun_out/tma_fill -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

using namespace istvt;

// grid = all SMs, one elected thread per CTA issues `iters` loads of a [rows x row_bytes] box into a ring of `stages`
// slots; a second warp waits on the full barriers and releases the slots (no MMA: pure fill rate).
__global__ void __launch_bounds__(64, 1)
fill_kernel(const __grid_constant__ CUtensorMap tm, int iters, int stages, int box_bytes, int rows_total, int box_rows,
            int k_boxes, unsigned long long* out_clk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + stages * box_bytes);
    uint64_t* empty = full + stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_mbar_init();
    }
    __syncthreads();
    const long long t0 = clock64();
    if (warp == 0 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        int row = (blockIdx.x * 977) % (rows_total - box_rows);
        for (int i = 0; i < iters; ++i) {
            mbar_wait_hot(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], box_bytes);
            tma_load_2d(base + stage * box_bytes, &tm, &full[stage], (i % k_boxes) * (box_bytes / box_rows / 2), row);
            if ((i % k_boxes) == k_boxes - 1) row = (row + box_rows * 131) % (rows_total - box_rows);
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int i = 0; i < iters; ++i) {
            mbar_wait_hot(&full[stage], phase);
            mbar_arrive(&empty[stage]);
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out_clk[blockIdx.x] = clock64() - t0;
}

static int g_grid = 148;
static void run(const char* name, void* buf, int64_t rows_total, int row_pitch_elems, int box_cols, int box_rows, int swz) {
    CUtensorMap tm;
    const int k_boxes = row_pitch_elems / box_cols > 0 ? row_pitch_elems / box_cols : 1;
    const uint64_t dims[2] = {static_cast<uint64_t>(row_pitch_elems), static_cast<uint64_t>(rows_total)};
    const uint64_t strides[1] = {static_cast<uint64_t>(row_pitch_elems) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(box_cols), static_cast<uint32_t>(box_rows)};
    if (encode_tmap(&tm, buf, ISTVT_BF16, 2, dims, strides, box, swz) != ISTVT_OK) { printf("%s: encode failed\n", name); return; }
    const int box_bytes = box_cols * 2 * box_rows;
    const int stages = 160 * 1024 / box_bytes > 16 ? 16 : 160 * 1024 / box_bytes;
    const int iters = 4000;
    unsigned long long* d_clk;
    cudaMalloc(&d_clk, 148 * 8);
    const int smem = stages * box_bytes + 1024 + 512;
    cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep)
        fill_kernel<<<g_grid, 64, smem>>>(tm, iters, stages, box_bytes, static_cast<int>(rows_total), box_rows, k_boxes, d_clk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    std::vector<unsigned long long> h(148);
    cudaMemcpy(h.data(), d_clk, 148 * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < g_grid; ++i) avg += static_cast<double>(h[i]);
    avg /= g_grid;
    printf("%-46s box %3d rows x %3d B = %5d B, %2d stages: %7.1f clk/box  %5.2f clk/row  %6.1f B/clk/SM\n", name, box_rows,
           box_cols * 2, box_bytes, stages, avg / iters, avg / iters / box_rows, box_bytes * iters / avg);
    cudaFree(d_clk);
}

int main(int argc, char** argv) {
    // argv[1] = CTAs (148: the memory system is shared; 8: what ONE SM's TMA unit can do), argv[2] = rows (162176:
    // 236 MB, HBM; 8192: L2-resident)
    g_grid = argc > 1 ? atoi(argv[1]) : 148;
    const int64_t rows = argc > 2 ? atoll(argv[2]) : 162176;
    printf("grid %d CTAs, %lld rows\n", g_grid, static_cast<long long>(rows));
    void* buf;
    cudaMalloc(&buf, rows * 768 * 2);
    cudaMemset(buf, 0, rows * 768 * 2);
    run("K=728 pitch (1456 B), 128 x 128 B, SW128", buf, rows, 728, 64, 128, 3);
    run("K=64 pitch (128 B, rows contiguous), SW128", buf, rows * 11, 64, 64, 128, 3);
    run("K=728 pitch, 128 x 64 B, SW64", buf, rows, 728, 32, 128, 2);
    run("K=32 pitch (64 B, contiguous), 128 x 64 B, SW64", buf, rows * 22, 32, 32, 128, 2);
    run("K=728 pitch, 256 x 128 B, SW128", buf, rows, 728, 64, 256, 3);
    run("K=728 pitch, 64 x 128 B, SW128", buf, rows, 728, 64, 64, 3);
    run("K=728 pitch, 128 x 128 B, no swizzle", buf, rows, 728, 64, 128, 0);
    run("K=2912 pitch (5824 B), 128 x 128 B, SW128", buf, rows / 4, 2912, 64, 128, 3);
    return 0;
}
