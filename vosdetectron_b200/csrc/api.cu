// Library-level entry points of libvosd_b200.so.
#include <atomic>
#include "common.cuh"

namespace vosd {
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace vosd

extern "C" const char* vosd_version(void) { return "vosd_b200 0.1.0 (sm_100a)"; }

extern "C" unsigned long long vosd_launch_count(void) {
    return vosd::g_launches.load(std::memory_order_relaxed);
}

extern "C" const char* vosd_status_string(int status) {
    switch (status) {
        case VOSD_OK: return "ok";
        case VOSD_ERR_BAD_SHAPE: return "bad shape";
        case VOSD_ERR_BAD_ARG: return "bad argument (null / misaligned pointer)";
        case VOSD_ERR_UNSUPPORTED: return "unsupported size (compiled-in limit)";
        case VOSD_ERR_WORKSPACE: return "workspace missing or too small";
        case VOSD_ERR_LAUNCH: return "CUDA launch error";
        default: return "unknown status";
    }
}

extern "C" int vosd_set_device(int device) {
    return cudaSetDevice(device) == cudaSuccess ? VOSD_OK : VOSD_ERR_BAD_ARG;
}
