// lc_order.h -- the order in which the reference schedules lane-change candidates.
//
// Engine::scheduleLaneChange (engine.cpp:792-794) std::sort()s the candidates by urgency, and every
// urgency is 1 (lanechange.cpp:184): the comparator never returns true, the sort is not stable, and
// the resulting order is whatever libstdc++'s introsort does to n indistinguishable elements -- a
// permutation that depends on n alone.  With a comparator that is always false, introsort
// (bits/stl_algo.h: __introsort_loop, __move_median_to_first, __unguarded_partition) reduces to:
// while a range is longer than 16, swap its first and middle element, then swap the pairs
// (first+1, last-1), (first+2, last-2), ... until they meet; recurse into the right part, continue
// with the left; the final insertion sort moves nothing.  This header replays that, so the device
// path can order its candidates without calling std::sort on the host.
// Not used by the engine yet (laneChange is rejected at load); pinned by tests/lc_order_probe.cpp
// against std::sort itself.
#pragma once
#include <utility>
#include <vector>

namespace cfb {

// perm[i] = input position of the element std::sort leaves at position i
inline void allEqualSortPermutation(int n, std::vector<int> &perm) {
    perm.resize(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    struct Range { int first, last; };
    std::vector<Range> todo;
    if (n > 0) todo.push_back({0, n});
    while (!todo.empty()) {
        Range r = todo.back();
        todo.pop_back();
        int first = r.first, last = r.last;
        while (last - first > 16) {
            const int mid = first + (last - first) / 2;
            std::swap(perm[first], perm[mid]);          // median of three indistinguishable elements -> the middle one
            int f = first + 1, l = last;
            for (;;) {                                  // unguarded partition around *first, comparator always false
                --l;
                if (!(f < l)) break;
                std::swap(perm[f], perm[l]);
                ++f;
            }
            todo.push_back({f, last});                  // the reference recurses right first; the ranges are disjoint, order is irrelevant
            last = f;
        }
    }
}

}  // namespace cfb
