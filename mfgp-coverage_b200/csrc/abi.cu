// Library-level entry points of libmfgp_b200: version string and per-thread CUDA error text.
#include <atomic>
#include <cstring>

#include "common.cuh"

namespace mfgp {
static thread_local char g_last_error[512] = "";
void set_last_error(const char* what, cudaError_t e) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace mfgp

extern "C" const char* mfgp_version(void) { return "mfgp_b200 0.1 (sm_100a, DMMA fp64)"; }
extern "C" const char* mfgp_last_error(void) { return mfgp::g_last_error; }
extern "C" int64_t mfgp_launch_count(void) { return mfgp::g_launches.load(std::memory_order_relaxed); }
