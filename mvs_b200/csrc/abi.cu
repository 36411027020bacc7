// Library-level entry points of include/mvs_b200.h (version, error string, launch counter).
#include "common.cuh"

namespace mvs {
std::string &last_error_ref()
{
    static thread_local std::string s;
    return s;
}
std::atomic<long long> g_launches{0};

int sm_count()
{
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
}  // namespace mvs

extern "C" int mvs_version(void) { return 100; }   // 0.1.0

extern "C" int mvs_sm(void)
{
#ifdef MVS_TARGET_SM
    return MVS_TARGET_SM;
#else
    return 100;
#endif
}

extern "C" const char *mvs_last_error(void) { return mvs::last_error_ref().c_str(); }

extern "C" int64_t mvs_launch_count(void) { return (int64_t)mvs::g_launches.load(); }
