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

// fp16 range guard (tc_common.cuh): one registered ring of 32-bit slots per device, handed out round robin
struct RangeRing {
    std::atomic<uint32_t *> base{nullptr};
    std::atomic<unsigned> n{0}, next{0};
};
static RangeRing g_range[64];
uint32_t *range_slot_next()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    RangeRing &r = g_range[dev];
    uint32_t *base = r.base.load(std::memory_order_acquire);
    const unsigned n = r.n.load(std::memory_order_relaxed);
    if (!base || !n) return nullptr;
    return base + r.next.fetch_add(1, std::memory_order_relaxed) % n;
}

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

extern "C" int side_tc_range_guard(void *status_words, int nwords)
{
    int dev = 0;
    SIDE_CUDA(cudaGetDevice(&dev));
    SIDE_REQUIRE(dev >= 0 && dev < 64, "side_tc_range_guard: device index out of range");
    side::RangeRing &r = side::g_range[dev];
    if (!status_words) {
        r.base.store(nullptr, std::memory_order_release);
        r.n.store(0);
        return SIDE_OK;
    }
    SIDE_REQUIRE(nwords >= 1, "side_tc_range_guard: nwords must be >= 1");
    SIDE_REQUIRE_DEV(status_words);
    r.base.store(nullptr, std::memory_order_release);
    r.n.store((unsigned)nwords);
    r.next.store(0);
    r.base.store(reinterpret_cast<uint32_t *>(status_words), std::memory_order_release);
    return SIDE_OK;
}

extern "C" long long side_launch_count(int reset)
{
    return reset ? side::g_launches.exchange(0) : side::g_launches.load();
}
