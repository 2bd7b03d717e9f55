// Host-side helpers shared by the kernel launchers and the runtime: status codes,
// last-error string, TMA tensor-map encoding through the driver entry point
// (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace sdod {

enum Status : int {
    kOk = 0,
    kInvalidArgument = -1,
    kCudaError = -2,
    kUnsupported = -3,
    kNoDevice = -4,
};

void set_last_error(const std::string& msg);
const char* last_error();
int fail(int code, const std::string& msg);          // records msg, returns code
int check_cuda(cudaError_t e, const char* what);     // kOk or kCudaError (+ message)
int check_launch(const char* what);                  // cudaGetLastError()

// Encode a tiled bf16 tensor map. dims/box are innermost-first; strides_bytes has rank-1
// entries (stride of dim i+1). swizzle128: CU_TENSOR_MAP_SWIZZLE_128B else NONE.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);
// General form: elem_bytes 2 (bf16) or 4 (fp32); swizzle_bytes 0 / 32 / 64 / 128.
int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

int device_sm_count();

}  // namespace sdod

#define SDOD_TRY(expr)                 \
    do {                               \
        int _s = (expr);               \
        if (_s != ::sdod::kOk) return _s; \
    } while (0)
