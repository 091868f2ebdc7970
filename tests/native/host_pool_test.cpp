// Stress test of the host-thread pool behind the pageable staging / complex128 widening of the host path
// (dc_sand_b200/csrc/host_pool.h): every part of every job runs exactly once, for 0 .. 7 workers and 1 .. 40 parts, jobs
// back to back (the case a late worker could get wrong).  Built and run by tests/test_host_logic.py, under
// ThreadSanitizer when the compiler offers it.
#include <atomic>
#include <cstdio>
#include <random>
#include <vector>

#include "host_pool.h"

int main() {
    std::mt19937 rng(7);
    for (int workers : {0, 1, 3, 7}) {
        ddch::HostPool pool(workers);
        for (int it = 0; it < 2000; ++it) {
            const int parts = 1 + (int)(rng() % 40);
            std::vector<int> hit((size_t)parts, 0);
            std::atomic<int> total{0};
            pool.run(parts, [&](int i) {
                hit[(size_t)i]++;
                total++;
            });
            if (total != parts) {
                std::printf("FAIL: %d of %d parts ran\n", (int)total, parts);
                return 1;
            }
            for (int i = 0; i < parts; ++i)
                if (hit[(size_t)i] != 1) {
                    std::printf("FAIL: part %d ran %d times\n", i, hit[(size_t)i]);
                    return 1;
                }
        }
    }
    std::printf("host pool OK\n");
    return 0;
}
