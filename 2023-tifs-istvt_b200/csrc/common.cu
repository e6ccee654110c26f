// Host-side plumbing shared by all kernels: error strings, launch counter, per-device SM count,
// TMA tensor-map encoding through the driver entry point (resolved at run time, so the library
// loads on a box without libcuda and fails with ISTVT_ERR_NO_DRIVER only when a TMA kernel is used).
#include "common.cuh"

#include <atomic>
#include <mutex>

namespace istvt {

static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static std::once_flag once;
    static EncodeTiledFn fn = nullptr;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    });
    return fn;
}

int encode_tmap(CUtensorMap* out, const void* base, int elem, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return ISTVT_ERR_NO_DRIVER;
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    const CUtensorMapDataType dt = elem == ISTVT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUtensorMapSwizzle sw = swizzle == 3   ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                 : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ISTVT_OK : ISTVT_ERR_TMAP;
}

}  // namespace istvt

extern "C" int istvt_abi_version(void) { return 1; }

extern "C" int64_t istvt_launch_count(void) { return istvt::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* istvt_error_string(int code) {
    switch (code) {
        case ISTVT_OK: return "ok";
        case ISTVT_ERR_INVALID_ARG: return "istvt: invalid argument (null pointer, bad size, alignment or dtype)";
        case ISTVT_ERR_NO_DRIVER: return "istvt: cuTensorMapEncodeTiled driver entry point unavailable";
        case ISTVT_ERR_TMAP: return "istvt: tensor-map encoding rejected the shape/strides";
        case ISTVT_ERR_UNSUPPORTED: return "istvt: configuration not supported by this kernel";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "istvt: unknown error code";
}
