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
// elem_strides (optional, rank entries): traversal stride per dimension; box[i] is then the extent of the traversed bounding box
// (ceil(box[i] / elem_strides[i]) elements are moved) — how a stride-2 convolution reads every second pixel.
int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes, const uint32_t* elem_strides = nullptr);

int device_sm_count();
bool pdl_enabled();   // SDOD_PDL=0 in the environment turns programmatic dependent launch off (A/B measurements)
// Per-thread switch for the launches that follow (plans decide per batch size: measured r1, PDL gains 2.9 % on the batch-2 UNet
// step, whose kernels are mostly single-wave, and costs 1.2 % on the batch-8 step).  Returns the previous value.
bool set_pdl_for_thread(bool on);

// Launch `kernel` so that it may begin (up to its griddepcontrol.wait) before the previous kernel on the stream has drained.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    // only grids that fit the chip in one go: measured r1, the attribute on multi-wave kernels cost the batch-8 step 2 %
    const bool small = static_cast<unsigned long long>(grid.x) * grid.y * grid.z <= 296ull;
    cfg.attrs = attr; cfg.numAttrs = (small && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace sdod

#define SDOD_TRY(expr)                 \
    do {                               \
        int _s = (expr);               \
        if (_s != ::sdod::kOk) return _s; \
    } while (0)
