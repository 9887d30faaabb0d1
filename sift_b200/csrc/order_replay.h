// Exact replay of the reference's cleanup sort (sift.cpp:37-42): std::sort(points, cmpByFilter) followed by
// keeping the leading unfiltered points.  The comparator (interestpoint.hpp:57-62) only looks at one bit, so
// libstdc++'s introsort does not sort anything among equal keys — but it is unstable, and the order in which the
// unfiltered points come out decides which earlier windows each descriptor sees.  Filtered points
// are dropped afterwards, so only the fate of the (few) unfiltered ones matters: this file simulates libstdc++'s
// std::sort (__introsort_loop with median-of-three + __unguarded_partition, depth limit 2*lg(n) with the
// heap-sort fallback, final insertion sort) on the sparse set of unfiltered positions.  Cost O(n) to set up plus
// roughly O(unfiltered * log n), instead of O(n log n) comparisons with unpredictable branches.
//
// Key: 0 = unfiltered (sorts first), 1 = filtered.  comp(a, b) == (a == 0 && b == 1).
// tests/test_abi_cpu.py compares the result with the real std::sort over many sizes and densities.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace siftgpu {

class SparseFilterSort {
   public:
    // zero_pos: ascending positions (indices into the n-element vector) of the unfiltered elements.
    // Returns, in post-sort order, the index into zero_pos of each unfiltered element.
    std::vector<uint32_t> run(uint32_t n, const std::vector<uint32_t>& zero_pos) {
        who_.assign(n, -1);
        zp_ = zero_pos;
        for (size_t i = 0; i < zp_.size(); ++i) who_[zp_[i]] = (int32_t)i;
        if (n > 1) {
            int lg = 0;
            for (uint32_t v = n; v > 1; v >>= 1) ++lg;  // std::__lg
            introsort_loop(0, (int64_t)n, 2 * lg, 0, zp_.size());
        }
        // __final_insertion_sort never lets an element pass an equal one: the unfiltered come out in position order
        std::vector<uint32_t> order(zp_.size());
        for (size_t i = 0; i < zp_.size(); ++i) order[i] = (uint32_t)who_[zp_[i]];
        return order;
    }

   private:
    std::vector<int32_t> who_;   // position -> unfiltered id, or -1 for a filtered element
    std::vector<uint32_t> zp_;   // positions of the unfiltered elements; every active segment's slice is sorted
    std::vector<uint32_t> tmp_, moved_;

    bool is_one(int64_t pos) const { return who_[(size_t)pos] < 0; }

    // zp_[za, zb) are the unfiltered positions inside [f, l)
    void introsort_loop(int64_t f, int64_t l, int depth, size_t za, size_t zb) {
        while (l - f > 16) {
            if (za == zb) return;  // only filtered elements: whatever happens here is unobservable
            if (depth == 0) {
                heap_fallback(f, l, za, zb);
                return;
            }
            --depth;
            const int64_t cut = partition_pivot(f, l, za, zb);
            const size_t zc = (size_t)(std::lower_bound(zp_.begin() + (long)za, zp_.begin() + (long)zb, (uint32_t)cut) - zp_.begin());
            introsort_loop(cut, l, depth, zc, zb);
            l = cut;
            zb = zc;
        }
    }

    // std::__partial_sort(first, last, last): materialise the segment and let libstdc++ do it
    void heap_fallback(int64_t f, int64_t l, size_t za, size_t zb) {
        std::vector<uint32_t> e((size_t)(l - f));
        for (int64_t p = f; p < l; ++p) e[(size_t)(p - f)] = who_[(size_t)p] < 0 ? 0x80000000u : (uint32_t)who_[(size_t)p];
        std::partial_sort(e.begin(), e.end(), e.end(), [](uint32_t a, uint32_t b) { return !(a >> 31) && (b >> 31); });
        size_t z = za;
        for (int64_t p = f; p < l; ++p) {
            const uint32_t v = e[(size_t)(p - f)];
            who_[(size_t)p] = (v >> 31) ? -1 : (int32_t)v;
            if (!(v >> 31)) zp_[z++] = (uint32_t)p;
        }
        (void)zb;
    }

    // contents of positions p and q trade places (p < q); keeps zp_[za, zb) sorted
    void swap_positions(int64_t p, int64_t q, size_t za, size_t zb) {
        const int32_t a = who_[(size_t)p], b = who_[(size_t)q];
        if (a < 0 && b < 0) return;
        who_[(size_t)p] = b;
        who_[(size_t)q] = a;
        if (a >= 0 && b >= 0) return;  // two unfiltered trade ids: the position set is unchanged
        auto beg = zp_.begin() + (long)za, end = zp_.begin() + (long)zb;
        if (a >= 0) {  // the unfiltered one moves up from p to q
            auto it = std::lower_bound(beg, end, (uint32_t)p);
            auto to = std::lower_bound(beg, end, (uint32_t)q);  // first element > q after removal sits here
            std::move(it + 1, to, it);
            *(to - 1) = (uint32_t)q;
        } else {       // the unfiltered one moves down from q to p
            auto it = std::lower_bound(beg, end, (uint32_t)q);
            auto to = std::lower_bound(beg, end, (uint32_t)p);
            std::move_backward(to, it, it + 1);
            *to = (uint32_t)p;
        }
    }

    // __unguarded_partition_pivot: __move_median_to_first(first, first+1, mid, last-1) then __unguarded_partition(first+1, last, first)
    int64_t partition_pivot(int64_t f, int64_t l, size_t za, size_t zb) {
        const int64_t mid = f + (l - f) / 2;
        const int a = is_one(f + 1), b = is_one(mid), c = is_one(l - 1);
        int64_t pick;  // see __move_median_to_first with comp(x, y) = (x == 0 && y == 1)
        if (a < b) pick = (b < c) ? mid : ((a < c) ? l - 1 : f + 1);
        else if (a < c) pick = f + 1;
        else if (b < c) pick = l - 1;
        else pick = mid;
        swap_positions(f, pick, za, zb);
        return is_one(f) ? partition_pivot_filtered(f, l, za, zb) : partition_pivot_unfiltered(f, l, za, zb);
    }

    // Pivot is a filtered element.  `first` skips unfiltered elements and stops at every filtered one; `last` steps down by
    // one per round and never skips.  Round t swaps the t-th filtered element from the left (position o_t) with position
    // l-t: an unfiltered element sitting at l-t drops to o_t.  Positions above `last` are filtered afterwards, so the scan
    // of `first` stops at l-t+1 at the latest.
    int64_t partition_pivot_filtered(int64_t f, int64_t l, size_t za, size_t zb) {
        const uint32_t* Z = zp_.data() + za;  // unfiltered positions in [f+1, l): the pivot at f is filtered
        const int64_t m = (int64_t)(zb - za);
        const int64_t ones = (l - f - 1) - m;
        // o(t) = position of the t-th filtered element of [f+1, l) = f + t + #{i : Z[i] - (f+1) - i < t}
        auto below = [&](int64_t t) {  // number of unfiltered elements before the t-th filtered one
            int64_t lo = 0, hi = m;
            while (lo < hi) {
                const int64_t md = (lo + hi) / 2;
                if ((int64_t)Z[md] - (f + 1) - md < t) lo = md + 1; else hi = md;
            }
            return lo;
        };
        auto goes_on = [&](int64_t t) { return t <= ones && f + t + below(t) < l - t; };
        int64_t lo = 1, hi = ones + 1;  // smallest t for which the round does not swap
        while (lo < hi) {
            const int64_t md = (lo + hi) / 2;
            if (goes_on(md)) lo = md + 1; else hi = md;
        }
        const int64_t T = lo;
        const int64_t cut = (T <= ones && f + T + below(T) == l - T) ? l - T : l - T + 1;
        // unfiltered elements at l-t, 1 <= t < T, drop to o(t); they are the top of the slice
        int64_t top = m;  // Z[top, m) move
        while (top > 0 && (int64_t)Z[top - 1] >= l - T + 1) --top;
        if (top < m) {
            moved_.clear();
            int64_t ip = 0;
            for (int64_t i = m - 1; i >= top; --i) {  // descending position = ascending t = ascending destination
                const int64_t p = Z[i], t = l - p;
                while (ip < top && (int64_t)Z[ip] - (f + 1) - ip < t) ++ip;
                const int64_t o = f + t + ip;
                who_[(size_t)o] = who_[(size_t)p];
                who_[(size_t)p] = -1;
                moved_.push_back((uint32_t)o);
            }
            tmp_.resize((size_t)m);
            std::merge(Z, Z + top, moved_.begin(), moved_.end(), tmp_.begin());
            std::copy(tmp_.begin(), tmp_.end(), zp_.begin() + (long)za);
        }
        return cut;
    }

    // Pivot is an unfiltered element.  `first` never skips (f+1, f+2, ...); `last` skips filtered elements, i.e. jumps from
    // one unfiltered element to the next lower one.  Round t swaps position f+t with the t-th unfiltered element from the top.
    int64_t partition_pivot_unfiltered(int64_t f, int64_t l, size_t za, size_t zb) {
        (void)l;
        uint32_t* Z = zp_.data() + za + 1;  // zp_[za] == f is the pivot
        const int64_t m = (int64_t)(zb - za) - 1;
        int64_t T = 1;
        tmp_.clear();  // top positions that keep an unfiltered element (the one swapped up), in descending order
        for (;; ++T) {
            if (T > m || (int64_t)Z[m - T] <= f + T) break;
            const int64_t P = f + T, Q = Z[m - T];
            const int32_t low = who_[(size_t)P];
            who_[(size_t)P] = who_[(size_t)Q];
            who_[(size_t)Q] = low;
            if (low >= 0) tmp_.push_back((uint32_t)Q);
        }
        if (T > 1) {
            // new position set: f+1 .. f+T-1, then the untouched ones >= f+T below the consumed top, then the kept top positions
            moved_.clear();
            for (int64_t t = 1; t < T; ++t) moved_.push_back((uint32_t)(f + t));
            for (int64_t i = 0; i < m - (T - 1); ++i)
                if ((int64_t)Z[i] >= f + T) moved_.push_back(Z[i]);
            for (size_t i = tmp_.size(); i-- > 0;) moved_.push_back(tmp_[i]);
            std::copy(moved_.begin(), moved_.end(), Z);
        }
        return f + T;
    }
};

}  // namespace siftgpu
