#include "../host_common.h"
#include "../launch_count.h"
#include "sdod_kernels.h"

namespace sdod {
std::atomic<unsigned long long> g_launch_count{0};
}

extern "C" {
SDOD_API const char* sdod_last_error(void) { return sdod::last_error(); }
SDOD_API int sdod_abi_version(void) { return 1; }
SDOD_API unsigned long long sdod_launch_count(void) { return sdod::g_launch_count.load(); }
}
