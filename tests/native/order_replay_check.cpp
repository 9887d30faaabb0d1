// Test helper (CPU): SparseFilterSort (own dense introsort by default; -DSIFT_ORDER_REPLAY_LIBSTDCXX_DENSE hands dense segments to
// libstdc++'s internal loop instead) against the real std::sort with the reference's comparator
// (interestpoint.hpp:57-62) over many sizes, densities and layouts.  Exit code 0 = every case agrees.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include "../../sift_b200/csrc/order_replay.h"

struct P { uint32_t id; bool filtered; };

int main() {
    std::mt19937 rng(12345);
    siftgpu::SparseFilterSort sorter;
    int cases = 0;
    for (uint32_t n : {1u, 2u, 15u, 16u, 17u, 33u, 100u, 257u, 1000u, 4097u, 20000u, 140000u}) {
        for (double density : {0.0, 0.001, 0.014, 0.1, 0.3, 0.5, 0.7, 0.95, 1.0}) {
            for (int layout = 0; layout < 4; ++layout) {
                std::vector<P> v(n);
                std::vector<uint32_t> zero_pos;
                std::uniform_real_distribution<double> u(0.0, 1.0);
                for (uint32_t i = 0; i < n; ++i) {
                    bool unf = u(rng) < density;
                    if (layout == 1) unf = unf && i < n / 2;           // everything unfiltered in the first half
                    if (layout == 2) unf = unf || (i % 97 == 0 && density > 0);   // a regular comb on top
                    if (layout == 3) unf = unf && (i / 64) % 2 == 0;   // bursts
                    v[i] = P{i, !unf};
                    if (unf) zero_pos.push_back(i);
                }
                std::sort(v.begin(), v.end(), [](const P& a, const P& b) { return !a.filtered && b.filtered; });
                const std::vector<uint32_t> got = sorter.run(n, zero_pos);   // indices into zero_pos, post-sort order
                if (got.size() != zero_pos.size()) { std::printf("size mismatch n=%u\n", n); return 1; }
                for (size_t i = 0; i < got.size(); ++i)
                    if (v[i].filtered || v[i].id != zero_pos[got[i]]) {
                        std::printf("mismatch n=%u density=%g layout=%d at %zu\n", n, density, layout, i);
                        return 1;
                    }
                ++cases;
            }
        }
    }
    std::printf("%d cases agree\n", cases);
    return 0;
}
