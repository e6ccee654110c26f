// Host-side helpers shared by all translation units: error codes, TMA descriptor encoding via the
// driver entry point (no link-time dependency on libcuda), per-device SM-count cache.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/istvt_b200.h"

namespace istvt {

#define ISTVT_CHECK_CUDA(expr)                                  \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) return static_cast<int>(_e);     \
    } while (0)

#define ISTVT_REQUIRE(cond)                          \
    do {                                             \
        if (!(cond)) return ISTVT_ERR_INVALID_ARG;   \
    } while (0)

// last launch error on this thread's device -> return code
inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? ISTVT_OK : static_cast<int>(e);
}

int sm_count();  // SMs of the current device (cached per device, thread-safe)
void count_launch();  // bump the process-wide launch counter (istvt_launch_count)

// Encode a tiled tensor map. dims/strides/box are innermost-first; strides has rank-1 entries (bytes).
// elem: 0 = bf16, 1 = fp32.  swizzle: 0 none, 1 32B, 2 64B, 3 128B.  Returns ISTVT_OK or an error code.
int encode_tmap(CUtensorMap* out, const void* base, int elem, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle);

}  // namespace istvt
