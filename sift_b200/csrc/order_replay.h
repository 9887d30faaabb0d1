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
// Which std::sort: GNU libstdc++'s (median of first+1 / mid / last-1 moved to first, unguarded Hoare partition, 16-element
// threshold, depth limit 2*lg n, heap sort below it) — the library the reference's shipped binary (GCC 7.1) and this build
// (GCC 13) both use; the algorithm has not changed from GCC 4.9 to 15.  Everything here is written out against that
// description and needs only public API (std::partial_sort for the heap fallback); with -DSIFT_ORDER_REPLAY_LIBSTDCXX_DENSE
// dense segments are handed to libstdc++'s own std::__introsort_loop instead (a cross-check used by the tests, slower).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace siftgpu {

class SparseFilterSort {
   public:
    // zero_pos[0, m): ascending positions (indices into the n-element vector) of the unfiltered elements.
    // Returns, in post-sort order, the index into zero_pos of each unfiltered element (valid until the next run).
    // State is two small parallel arrays (position, id) kept sorted by position: nothing of size n is ever touched.
    const std::vector<uint32_t>& run(uint32_t n, const uint32_t* zero_pos, size_t m) {
        zp_.assign(zero_pos, zero_pos + m);
        id_.resize(m);
        for (size_t i = 0; i < m; ++i) id_[i] = (uint32_t)i;
        if (n > 1) {
            int lg = 0;
            for (uint32_t v = n; v > 1; v >>= 1) ++lg;  // std::__lg
            introsort_loop(0, (int64_t)n, 2 * lg, 0, m);
        }
        // __final_insertion_sort never lets an element pass an equal one: the unfiltered come out in position order
        return id_;
    }
    const std::vector<uint32_t>& run(uint32_t n, const std::vector<uint32_t>& zero_pos) { return run(n, zero_pos.data(), zero_pos.size()); }

    // A segment is materialised once this fraction of it is unfiltered (dense_num / dense_den; measured optimum).
    int dense_num = 1, dense_den = 2;

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
            if ((zb - za) * (size_t)dense_den >= (size_t)(l - f) * (size_t)dense_num) {  // cheaper to sort the real thing from here on
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

    static bool word_less(uint32_t a, uint32_t b) { return !(a >> 31) && (b >> 31); }

    // std::__partial_sort(first, last, last): materialise the segment and let the library's heap sort do it
    void heap_fallback(int64_t f, int64_t l, size_t za, size_t zb) {
        materialise(f, l, za, zb);
        std::partial_sort(seg_.begin(), seg_.end(), seg_.end(), word_less);
        collect(f, l, za, zb);
    }

    // segment [f, l) as words: bit 31 = filtered, low bits = id of the unfiltered element sitting there
    void materialise(int64_t f, int64_t l, size_t za, size_t zb) {
        seg_.assign((size_t)(l - f), 0x80000000u);
        for (size_t i = za; i < zb; ++i) seg_[(size_t)((int64_t)zp_[i] - f)] = id_[i];
    }
    // back to (position, id) pairs in position order; branch-free compaction through the scratch arrays
    void collect(int64_t f, int64_t l, size_t za, size_t zb) {
        const size_t len = (size_t)(l - f);
        tp_.resize(len + 1);
        ti_.resize(len + 1);
        size_t z = 0;
        for (size_t p = 0; p < len; ++p) {
            const uint32_t v = seg_[p];
            tp_[z] = (uint32_t)(f + (int64_t)p);
            ti_[z] = v;
            z += (v >> 31) ^ 1u;
        }
        std::copy(tp_.begin(), tp_.begin() + (long)(zb - za), zp_.begin() + (long)za);
        std::copy(ti_.begin(), ti_.begin() + (long)(zb - za), id_.begin() + (long)za);
    }

    void dense_segment(int64_t f, int64_t l, int depth, size_t za, size_t zb) {
        materialise(f, l, za, zb);
#if defined(SIFT_ORDER_REPLAY_LIBSTDCXX_DENSE)
        std::__introsort_loop(seg_.begin(), seg_.end(), (long)depth, __gnu_cxx::__ops::__iter_comp_iter(word_less));
#else
        dense_introsort(seg_.data(), seg_.data() + seg_.size(), depth, zb - za);
#endif
        collect(f, l, za, zb);
    }

    // libstdc++'s __introsort_loop for this comparator (key = bit 31; comp(a, b) = key(a) < key(b)), written out:
    // __move_median_to_first(first, first + 1, mid, last - 1), __unguarded_partition(first + 1, last, first), recursion on the
    // right part, iteration on the left, 16-element threshold, heap sort when the depth budget is used up.  `zeros` = unfiltered
    // elements in the segment: a segment without any is left alone (nothing that happens to it can be observed).
    //
    // The two partition loops are restated without data-dependent branches.  With a filtered pivot `lo` stops at every filtered
    // element and `hi` steps down by one per round, so round t swaps the t-th filtered element from the left with position
    // last-t; with an unfiltered pivot `lo` steps up by one per round and `hi` jumps from one unfiltered element to the next
    // lower one, so round t swaps position first+t with the t-th unfiltered element from the right.  Either way the positions
    // the scanning pointer stops at are listed block by block with a branch-free compaction, and the swap loop has one
    // (predictable) exit test.  Reading a block's contents before the swaps of that block is safe: the swaps only write behind
    // the scanning pointer or beyond the other pointer, where the pointers have crossed by the time the scan gets there.
    static void dense_introsort(uint32_t* first, uint32_t* last, int depth, size_t zeros) {
        constexpr int kBlock = 64;
        uint32_t* stops[kBlock + 1];
        while (last - first > 16) {
            if (zeros == 0) return;
            if (depth == 0) {
                std::partial_sort(first, last, last, word_less);
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
            uint32_t* cut = nullptr;
            if (zeros == (size_t)(last - first)) {
                // nothing but unfiltered elements (the pivot is `mid`): neither pointer ever skips, round t swaps first+t with
                // last-t until they meet — [first+1, last) is reversed, and both halves go on being shuffled like this
                uint32_t *lo = first + 1, *hi = last - 1;
                for (; lo < hi; ++lo, --hi) std::swap(*lo, *hi);
                cut = lo;
                const size_t left = (size_t)(cut - first);
                dense_introsort(cut, last, depth, zeros - left);
                last = cut;
                zeros = left;
            } else if (*first >> 31) {
                // pivot filtered.  hi_prev = last - (t - 1); positions >= hi_prev hold filtered elements from round 2 on.
                uint32_t* hi_prev = last;
                uint32_t* blk = first + 1;
                while (!cut) {
                    if (blk >= hi_prev) { cut = hi_prev; break; }  // nothing filtered left below hi_prev: lo runs into it
                    uint32_t* be = blk + kBlock < last ? blk + kBlock : last;
                    int k = 0;
                    for (uint32_t* q = blk; q < be; ++q) { stops[k] = q; k += (int)(*q >> 31); }
                    for (int j = 0; j < k; ++j) {
                        uint32_t* lo = stops[j] < hi_prev ? stops[j] : hi_prev;
                        uint32_t* hi = hi_prev - 1;
                        if (!(lo < hi)) { cut = lo; break; }
                        std::swap(*lo, *hi);
                        hi_prev = hi;
                    }
                    blk = be;
                }
                // [cut, last) holds filtered elements only
                last = cut;
            } else {
                // pivot unfiltered.  lo_prev = first + (t - 1); positions <= lo_prev hold unfiltered elements.
                uint32_t* lo_prev = first;
                uint32_t* blk = last;  // the block below `blk` is listed next
                while (!cut) {
                    if (blk <= lo_prev + 1) { cut = lo_prev + 1; break; }  // hi runs down into the unfiltered prefix
                    uint32_t* bs = (blk - first) > kBlock ? blk - kBlock : first + 1;
                    int k = 0;
                    for (uint32_t* q = blk - 1; q >= bs; --q) { stops[k] = q; k += (int)((*q >> 31) ^ 1u); }
                    for (int j = 0; j < k; ++j) {
                        uint32_t* lo = lo_prev + 1;
                        uint32_t* hi = stops[j] > lo_prev ? stops[j] : lo_prev;
                        if (!(lo < hi)) { cut = lo; break; }
                        std::swap(*lo, *hi);
                        lo_prev = lo;
                    }
                    blk = bs;
                }
                // [first, cut) holds unfiltered elements only
                const size_t left = (size_t)(cut - first);
                dense_introsort(cut, last, depth, zeros - left);
                last = cut;
                zeros = left;
            }
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
            // Merge the untouched bottom of the slice (ascending; element i goes before the dropped element of round t iff fewer
            // than t filtered elements precede it) with the dropped elements (descending position = ascending t = ascending
            // destination f + t + #bottom elements before it), ids riding along.  Everything below the first destination keeps
            // its place.  The loop selects instead of branching: which side comes next is a coin flip.
            const int64_t x0 = std::min<int64_t>(below(l - (int64_t)Z[m - 1]), top);
            tp_.resize((size_t)(m - x0));
            ti_.resize((size_t)(m - x0));
            uint32_t* TP = tp_.data();
            uint32_t* TI = ti_.data();
            const int64_t base = f + 1;
            int64_t ia = x0, ib = m - 1, o = 0;
            while (ia < top && ib >= top) {
                const int64_t za_ = Z[ia], t = l - (int64_t)Z[ib];
                const bool stay_first = (za_ - base - ia) < t;
                TP[o] = stay_first ? (uint32_t)za_ : (uint32_t)(f + t + ia);
                TI[o] = stay_first ? I[ia] : I[ib];
                ++o;
                ia += stay_first ? 1 : 0;
                ib -= stay_first ? 0 : 1;
            }
            for (; ia < top; ++ia, ++o) { TP[o] = Z[ia]; TI[o] = I[ia]; }
            for (; ib >= top; --ib, ++o) { TP[o] = (uint32_t)(f + (l - (int64_t)Z[ib]) + top); TI[o] = I[ib]; }
            std::copy(TP, TP + (m - x0), Z + x0);
            std::copy(TI, TI + (m - x0), I + x0);
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
