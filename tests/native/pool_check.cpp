// Test helper (CPU): the worker pool of the host pipeline (sift_b200/csrc/pool.h) under the usage pattern of sift_gpu_run —
// two pools, one job each in flight, the next job started right after the previous one was joined, jobs of every size
// including empty ones, repeated end() calls (the error-path quiesce).  Every item must run exactly once and nothing may
// hang: a worker that is still leaving job k when job k+1 starts must not draw from it.  Exit code 0 = all counts agree.
// Built with -fsanitize=thread by the test as well.
#include <atomic>
#include <cstdio>
#include <vector>

#include "../../sift_b200/csrc/pool.h"

int main(int argc, char** argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 20000;
    siftgpu::Pool replay(3), pack(3);
    std::vector<int> a(4096), b(4096);
    std::atomic<long> sum{0};
    long expect = 0;
    int pack_n = 37;
    pack.begin(pack_n, [&](int i) { b[(size_t)i] += 1; sum += i; });
    for (int i = 0; i < pack_n; ++i) expect += i;
    for (int k = 0; k < rounds; ++k) {
        const int nb = 1 + (k * 7) % 192;
        std::fill(a.begin(), a.begin() + nb, 0);
        replay.begin(nb, [&](int i) { a[(size_t)i] += 1; sum += 1; });
        expect += nb;
        // every third round the rest of the pack job is cancelled (the split upload): items below the returned count ran exactly
        // once, the others not at all
        int ran = pack_n;
        if (k % 3 == 1) {
            ran = pack.cancel_rest();
            if (ran < 0 || ran > pack_n) { std::printf("cancel_rest returned %d of %d\n", ran, pack_n); return 1; }
            expect -= 2L * (pack_n - ran);
        }
        pack.end();
        for (int i = 0; i < pack_n; ++i)
            if (b[(size_t)i] != (i < ran ? 1 : 0)) { std::printf("pack item %d of round %d ran %d times (%d started)\n", i, k, b[(size_t)i], ran); return 1; }
        replay.end();
        for (int i = 0; i < nb; ++i)
            if (a[(size_t)i] != 1) { std::printf("replay item %d of round %d ran %d times\n", i, k, a[(size_t)i]); return 1; }
        pack_n = (k % 5 == 0) ? 0 : 1 + (k * 13) % 1536;
        std::fill(b.begin(), b.end(), 0);
        if (pack_n) {
            pack.begin(pack_n, [&](int i) { b[(size_t)i] += 1; sum += 2; });
            expect += 2L * pack_n;
        }
        if (k % 1000 == 0) { pack.end(); pack.end(); pack_n = 0; }
    }
    pack.end();
    std::printf("%s: sum %ld expect %ld\n", sum.load() == expect ? "pool ok" : "MISMATCH", sum.load(), expect);
    return sum.load() == expect ? 0 : 1;
}
