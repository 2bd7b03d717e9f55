// Process-wide count of kernels launched by this library (bench.py reports it as gpu_launches).
#pragma once
#include <atomic>
namespace sdod {
extern std::atomic<unsigned long long> g_launch_count;
inline void count_launch(unsigned long long n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }
}  // namespace sdod
