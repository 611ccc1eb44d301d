// api.cu -- library-wide pieces of the C ABI: error reporting, pointer checks, launch counter.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace side {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool is_device_ptr(const void *p)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

__global__ void probe_kernel(int *out) { *out = 100; }

}  // namespace side

extern "C" int side_abi_version(void) { return SIDE_ABI_VERSION; }

extern "C" const char *side_last_error(void) { return side::g_err; }

extern "C" int side_device_ok(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return 0;
    }
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, (const void *)side::probe_kernel) != cudaSuccess) {
        cudaGetLastError();
        return 0;  // no sm_100a image for this device
    }
    return 1;
}

extern "C" long long side_launch_count(int reset)
{
    return reset ? side::g_launches.exchange(0) : side::g_launches.load();
}
