// Exact replay of the reference's cleanup sort (sift.cpp:37-42): std::sort(points, cmpByFilter) followed by
// keeping the leading unfiltered points.  The comparator (interestpoint.hpp:57-62) only looks at one bit, so
// libstdc++'s introsort does not sort anything among equal keys — but it is unstable, and the order in which the
// unfiltered points come out decides which earlier windows each descriptor sees.  Filtered points
// are dropped afterwards, so only the fate of the (few) unfiltered ones matters: this file simulates libstdc++'s
// std::sort (__introsort_loop with median-of-three + __unguarded_partition, depth limit 2*lg(n) with the
// heap-sort fallback, final insertion sort) on the sparse set of unfiltered positions.  Cost roughly
// O(unfiltered * log n) on two small arrays, instead of O(n log n) comparisons with unpredictable branches.
//
// Key: 0 = unfiltered (sorts first), 1 = filtered.  comp(a, b) == (a == 0 && b == 1).
// tests/test_abi_cpu.py compares the result with the real std::sort over many sizes and densities.
//
// Which std::sort: the dense hand-off below calls libstdc++'s own std::__introsort_loop, an internal whose signature has
// been stable from GCC 4.9 to 15 (checked range below); outside that range, or on another standard library, every segment
// takes the sparse simulation, which implements the same algorithm (libstdc++'s: median of first+1 / mid / last-1 moved to
// first, unguarded Hoare partition, 16-element threshold, depth limit 2*lg n) and gives the same permutation, only slower.
// The order contract is therefore "GNU libstdc++'s introsort" — the library the reference's shipped binary (GCC 7.1) and
// this build (GCC 13) both use.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#if defined(__GLIBCXX__) && defined(_GLIBCXX_RELEASE) && _GLIBCXX_RELEASE >= 5 && _GLIBCXX_RELEASE <= 15
#define SIFT_ORDER_REPLAY_LIBSTDCXX 1
#endif

namespace siftgpu {

class SparseFilterSort {
   public:
    // zero_pos: ascending positions (indices into the n-element vector) of the unfiltered elements.
    // Returns, in post-sort order, the index into zero_pos of each unfiltered element.
    // State is two small parallel arrays (position, id) kept sorted by position: nothing of size n is ever touched.
    std::vector<uint32_t> run(uint32_t n, const std::vector<uint32_t>& zero_pos) {
        zp_ = zero_pos;
        id_.resize(zp_.size());
        for (size_t i = 0; i < zp_.size(); ++i) id_[i] = (uint32_t)i;
        if (n > 1) {
            int lg = 0;
            for (uint32_t v = n; v > 1; v >>= 1) ++lg;  // std::__lg
            introsort_loop(0, (int64_t)n, 2 * lg, 0, zp_.size());
        }
        // __final_insertion_sort never lets an element pass an equal one: the unfiltered come out in position order
        return id_;
    }

   private:
    std::vector<uint32_t> zp_;   // positions of the unfiltered elements, ascending; every active segment owns a slice
    std::vector<uint32_t> id_;   // id_[i] = which unfiltered element sits at zp_[i]
    std::vector<uint32_t> tp_, ti_, mp_, mi_, seg_;

    // index of `pos` in zp_[za, zb), or -1 when a filtered element sits there
    int64_t find(int64_t pos, size_t za, size_t zb) const {
        const auto beg = zp_.begin() + (long)za, end = zp_.begin() + (long)zb;
        const auto it = std::lower_bound(beg, end, (uint32_t)pos);
        return (it != end && *it == (uint32_t)pos) ? (int64_t)(it - zp_.begin()) : -1;
    }
    bool is_one(int64_t pos, size_t za, size_t zb) const { return find(pos, za, zb) < 0; }

    // zp_[za, zb) are the unfiltered positions inside [f, l)
    void introsort_loop(int64_t f, int64_t l, int depth, size_t za, size_t zb) {
        while (l - f > 16) {
            if (za == zb) return;  // only filtered elements: whatever happens here is unobservable
            if (depth == 0) {
                heap_fallback(f, l, za, zb);
                return;
            }
            if ((zb - za) * 2 >= (size_t)(l - f)) {  // half unfiltered: cheaper to let libstdc++ run on the real thing
                dense_segment(f, l, depth, za, zb);
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
        std::vector<uint32_t> e((size_t)(l - f), 0x80000000u);
        for (size_t i = za; i < zb; ++i) e[(size_t)((int64_t)zp_[i] - f)] = id_[i];
        std::partial_sort(e.begin(), e.end(), e.end(), [](uint32_t a, uint32_t b) { return !(a >> 31) && (b >> 31); });
        size_t z = za;
        for (int64_t p = f; p < l; ++p) {
            const uint32_t v = e[(size_t)(p - f)];
            if (!(v >> 31)) { zp_[z] = (uint32_t)p; id_[z] = v; ++z; }
        }
    }

    // A segment in which at least half of the elements is unfiltered is materialised as (filtered bit | id) words and sorted
    // for real.  With libstdc++ (checked release range above) by its own std::__introsort_loop with the depth budget left at
    // this point of the recursion — the same code std::sort would be running here; elsewhere by the restatement below
    // (SIFT_ORDER_REPLAY_OWN_DENSE selects it explicitly; tests compare both with std::sort).
    void dense_segment(int64_t f, int64_t l, int depth, size_t za, size_t zb) {
        seg_.assign((size_t)(l - f), 0x80000000u);
        for (size_t i = za; i < zb; ++i) seg_[(size_t)((int64_t)zp_[i] - f)] = id_[i];
#if defined(SIFT_ORDER_REPLAY_LIBSTDCXX) && !defined(SIFT_ORDER_REPLAY_OWN_DENSE)
        auto comp = [](uint32_t a, uint32_t b) { return !(a >> 31) && (b >> 31); };
        std::__introsort_loop(seg_.begin(), seg_.end(), (long)depth, __gnu_cxx::__ops::__iter_comp_iter(comp));
#else
        dense_introsort(seg_.data(), seg_.data() + seg_.size(), depth);
#endif
        size_t z = za;
        for (int64_t p = f; p < l; ++p) {
            const uint32_t v = seg_[(size_t)(p - f)];
            if (!(v >> 31)) { zp_[z] = (uint32_t)p; id_[z] = v; ++z; }
        }
    }

    // libstdc++'s __introsort_loop for this comparator (key = bit 31; comp(a, b) = key(a) < key(b)), written out:
    // __move_median_to_first(first, first + 1, mid, last - 1), __unguarded_partition(first + 1, last, first), recursion on the
    // right part, iteration on the left, 16-element threshold, heap sort when the depth budget is used up.
    static void dense_introsort(uint32_t* first, uint32_t* last, int depth) {
        while (last - first > 16) {
            if (depth == 0) {
                std::partial_sort(first, last, last, [](uint32_t a, uint32_t b) { return !(a >> 31) && (b >> 31); });
                return;
            }
            --depth;
            uint32_t* mid = first + (last - first) / 2;
            {
                uint32_t *a = first + 1, *b = mid, *c = last - 1;
                const uint32_t ka = *a >> 31, kb = *b >> 31, kc = *c >> 31;
                uint32_t* pick;
                if (ka < kb) pick = (kb < kc) ? b : ((ka < kc) ? c : a);
                else if (ka < kc) pick = a;
                else if (kb < kc) pick = c;
                else pick = b;
                std::swap(*first, *pick);
            }
            uint32_t *lo = first + 1, *hi = last;
            if (*first >> 31) {
                // pivot filtered: `lo` runs over unfiltered elements and stops at every filtered one, `hi` steps down by one
                for (;;) {
                    while (!(*lo >> 31)) ++lo;
                    --hi;
                    if (!(lo < hi)) break;
                    std::swap(*lo, *hi);
                    ++lo;
                }
            } else {
                // pivot unfiltered: `lo` never skips, `hi` skips filtered elements
                for (;;) {
                    --hi;
                    while (*hi >> 31) --hi;
                    if (!(lo < hi)) break;
                    std::swap(*lo, *hi);
                    ++lo;
                }
            }
            dense_introsort(lo, last, depth);
            last = lo;
        }
    }

    // contents of positions p and q trade places (p < q); keeps zp_[za, zb) sorted
    void swap_positions(int64_t p, int64_t q, size_t za, size_t zb) {
        const int64_t ip = find(p, za, zb), iq = find(q, za, zb);
        if (ip < 0 && iq < 0) return;
        if (ip >= 0 && iq >= 0) { std::swap(id_[(size_t)ip], id_[(size_t)iq]); return; }
        const auto beg = zp_.begin() + (long)za, end = zp_.begin() + (long)zb;
        if (ip >= 0) {  // the unfiltered one moves up from p to q
            const size_t to = (size_t)(std::lower_bound(beg, end, (uint32_t)q) - zp_.begin());  // first position > q
            const uint32_t who = id_[(size_t)ip];
            std::move(zp_.begin() + ip + 1, zp_.begin() + (long)to, zp_.begin() + ip);
            std::move(id_.begin() + ip + 1, id_.begin() + (long)to, id_.begin() + ip);
            zp_[to - 1] = (uint32_t)q;
            id_[to - 1] = who;
        } else {        // the unfiltered one moves down from q to p
            const size_t to = (size_t)(std::lower_bound(beg, end, (uint32_t)p) - zp_.begin());  // first position > p
            const uint32_t who = id_[(size_t)iq];
            std::move_backward(zp_.begin() + (long)to, zp_.begin() + iq, zp_.begin() + iq + 1);
            std::move_backward(id_.begin() + (long)to, id_.begin() + iq, id_.begin() + iq + 1);
            zp_[to] = (uint32_t)p;
            id_[to] = who;
        }
    }

    // __unguarded_partition_pivot: __move_median_to_first(first, first+1, mid, last-1) then __unguarded_partition(first+1, last, first)
    int64_t partition_pivot(int64_t f, int64_t l, size_t za, size_t zb) {
        const int64_t mid = f + (l - f) / 2;
        const int a = is_one(f + 1, za, zb), b = is_one(mid, za, zb), c = is_one(l - 1, za, zb);
        int64_t pick;  // see __move_median_to_first with comp(x, y) = (x == 0 && y == 1)
        if (a < b) pick = (b < c) ? mid : ((a < c) ? l - 1 : f + 1);
        else if (a < c) pick = f + 1;
        else if (b < c) pick = l - 1;
        else pick = mid;
        swap_positions(f, pick, za, zb);
        return is_one(f, za, zb) ? partition_pivot_filtered(f, l, za, zb) : partition_pivot_unfiltered(f, l, za, zb);
    }

    // Pivot is a filtered element.  `first` skips unfiltered elements and stops at every filtered one; `last` steps down by
    // one per round and never skips.  Round t swaps the t-th filtered element from the left (position o_t) with position
    // l-t: an unfiltered element sitting at l-t drops to o_t.  Positions above `last` are filtered afterwards, so the scan
    // of `first` stops at l-t+1 at the latest.
    int64_t partition_pivot_filtered(int64_t f, int64_t l, size_t za, size_t zb) {
        uint32_t* Z = zp_.data() + za;  // unfiltered positions in [f+1, l): the pivot at f is filtered
        uint32_t* I = id_.data() + za;
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
            // Merge the untouched bottom of the slice with the dropped elements (both ascending), ids riding along.  Everything
            // below the first destination keeps its place in the arrays: after a few levels that is a dense prefix holding
            // most of the unfiltered elements, so the work per level shrinks with the part that is still sparse.
            int64_t ip = std::min<int64_t>(below(l - (int64_t)Z[m - 1]), top);
            const int64_t x0 = ip;
            tp_.resize((size_t)(m - x0));
            ti_.resize((size_t)(m - x0));
            int64_t x = x0, o = 0;
            for (int64_t i = m - 1; i >= top; --i) {  // descending position = ascending t = ascending destination
                const int64_t p = Z[i], t = l - p;
                while (ip < top && (int64_t)Z[ip] - (f + 1) - ip < t) ++ip;
                for (; x < ip; ++x, ++o) { tp_[(size_t)o] = Z[x]; ti_[(size_t)o] = I[x]; }
                tp_[(size_t)o] = (uint32_t)(f + t + ip);
                ti_[(size_t)o] = I[i];
                ++o;
            }
            for (; x < top; ++x, ++o) { tp_[(size_t)o] = Z[x]; ti_[(size_t)o] = I[x]; }
            std::copy(tp_.begin(), tp_.begin() + (long)(m - x0), Z + x0);
            std::copy(ti_.begin(), ti_.begin() + (long)(m - x0), I + x0);
        }
        return cut;
    }

    // Pivot is an unfiltered element.  `first` never skips (f+1, f+2, ...); `last` skips filtered elements, i.e. jumps from
    // one unfiltered element to the next lower one.  Round t swaps position f+t with the t-th unfiltered element from the top.
    int64_t partition_pivot_unfiltered(int64_t f, int64_t l, size_t za, size_t zb) {
        (void)l;
        uint32_t* Z = zp_.data() + za + 1;  // zp_[za] == f is the pivot
        uint32_t* I = id_.data() + za + 1;
        const int64_t m = (int64_t)(zb - za) - 1;
        int64_t T = 1, j = 0;
        mp_.clear();  // first the ids that land on f+1, f+2, ...
        mi_.clear();
        tp_.clear();  // top positions that keep an unfiltered element (the one swapped up), in descending order
        ti_.clear();
        for (;; ++T) {
            if (T > m || (int64_t)Z[m - T] <= f + T) break;
            const int64_t P = f + T, Q = Z[m - T];
            while (j < m && (int64_t)Z[j] < P) ++j;
            mi_.push_back(I[m - T]);  // the t-th unfiltered element from the top lands on f+t
            if (j < m && (int64_t)Z[j] == P) {  // an unfiltered element sat at f+t: it goes up to Q
                tp_.push_back((uint32_t)Q);
                ti_.push_back(I[j]);
            }
        }
        if (T > 1) {
            // new position set: f+1 .. f+T-1, then the untouched ones >= f+T below the consumed top, then the kept top positions
            for (int64_t t = 1; t < T; ++t) mp_.push_back((uint32_t)(f + t));
            for (int64_t i = 0; i < m - (T - 1); ++i)
                if ((int64_t)Z[i] >= f + T) { mp_.push_back(Z[i]); mi_.push_back(I[i]); }
            for (size_t i = tp_.size(); i-- > 0;) { mp_.push_back(tp_[i]); mi_.push_back(ti_[i]); }
            std::copy(mp_.begin(), mp_.end(), Z);
            std::copy(mi_.begin(), mi_.end(), I);
        }
        return f + T;
    }
};

}  // namespace siftgpu
