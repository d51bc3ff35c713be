// Host-side unit test of vpc_host::CopyPool (csrc/host/staging.hpp): every index of a job runs exactly once whatever the pool size,
// the number of workers the job admits, and the rhythm of the jobs (back to back, or after the workers have gone to sleep).
// No GPU: only the pool is instantiated.  Build: g++ -O2 -std=c++17 -pthread -I<cuda include>.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "../../vtkcloudpoint_b200/csrc/host/staging.hpp"

int main() {
  int failures = 0;
  for (int pool_size : {0, 1, 3, 8}) {
    vpc_host::CopyPool pool(pool_size);
    unsigned long long rng = 0x9e3779b97f4a7c15ull + (unsigned)pool_size;
    for (int job = 0; job < 3000; ++job) {
      rng = rng * 6364136223846793005ull + 1442695040888963407ull;
      const size_t n = (size_t)((rng >> 33) % 70);                    // 0 .. 69 items
      const int limit = (int)((rng >> 20) % 10);                      // 0 .. 9 admitted workers (0: the caller alone)
      std::vector<std::atomic<int>> hits(n);
      for (auto& h : hits) h.store(0);
      std::atomic<long long> sum{0};
      pool.parallel_for(n, [&](size_t i) { hits[i].fetch_add(1); sum.fetch_add((long long)i + 1); }, limit);
      long long want = 0;
      for (size_t i = 0; i < n; ++i) { want += (long long)i + 1; if (hits[i].load() != 1) ++failures; }
      if (sum.load() != want) ++failures;
      if (job % 500 == 499) std::this_thread::sleep_for(std::chrono::milliseconds(20));   // let the workers fall asleep
    }
  }
  if (failures) { std::printf("FAILED: %d\n", failures); return 1; }
  std::printf("copy pool ok\n");
  return 0;
}
