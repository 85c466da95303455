#include "common.h"

namespace se2gpu {
thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};
}  // namespace se2gpu

extern "C" {

int se2gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* se2gpu_last_error(void) { return se2gpu::g_last_error.c_str(); }

unsigned long long se2gpu_launch_count(void) { return se2gpu::g_launches.load(); }

}  // extern "C"
