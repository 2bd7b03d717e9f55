#include "host_common.h"

#include <cudaTypedefs.h>
#include <cstdlib>
#include <mutex>

namespace sdod {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* last_error() { return g_last_error.c_str(); }

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return kOk;
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return kCudaError;
}

int check_launch(const char* what) { return check_cuda(cudaGetLastError(), what); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
    return encode_tmap(out, base, 2, rank, dims, strides_bytes, box, swizzle128 ? 128 : 0);
}

int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes, const uint32_t* elem_strides) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(kNoDevice, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = elem_strides ? elem_strides[i] : 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(kInvalidArgument, "TMA base must be 16-byte aligned");
    for (int i = 0; i + 1 < rank; ++i)
        if (gstr[i] % 16 != 0) return fail(kInvalidArgument, "TMA strides must be multiples of 16 bytes");
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                    const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(kCudaError, "cuTensorMapEncodeTiled failed, CUresult=" + std::to_string(static_cast<int>(r)));
    return kOk;
}

static thread_local bool g_pdl_thread = true;
bool set_pdl_for_thread(bool on) {
    const bool prev = g_pdl_thread;
    g_pdl_thread = on;
    return prev;
}
bool pdl_enabled() {
    static const bool on = [] { const char* e = std::getenv("SDOD_PDL"); return !e || std::atoi(e) != 0; }();
    return on && g_pdl_thread;
}

int device_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace sdod
