// CPU probe: cfb::allEqualSortPermutation (cityflow_b200/csrc/lc_order.h) against what std::sort of
// this toolchain's libstdc++ really does to n elements under the reference's comparator
// (engine.cpp:793: a->urgency > b->urgency, all urgencies 1).
#include <algorithm>
#include <cstdio>
#include <vector>
#include "lc_order.h"

struct Cand { int urgency, id; };

int main() {
    std::vector<int> perm;
    int checked = 0;
    for (int n = 0; n <= 6000; n += (n < 600 ? 1 : 37)) {
        std::vector<Cand> v(n);
        for (int i = 0; i < n; ++i) v[i] = Cand{1, i};
        std::sort(v.begin(), v.end(), [](const Cand &a, const Cand &b) { return a.urgency > b.urgency; });
        cfb::allEqualSortPermutation(n, perm);
        for (int i = 0; i < n; ++i)
            if (v[i].id != perm[i]) { printf("FAIL n=%d position %d: std::sort %d replay %d\n", n, i, v[i].id, perm[i]); return 1; }
        ++checked;
    }
    bool moved = false;
    cfb::allEqualSortPermutation(40, perm);
    for (int i = 0; i < 40; ++i) moved |= perm[i] != i;
    if (!moved) { printf("FAIL: n=40 should not be the identity\n"); return 1; }
    printf("OK %d sizes\n", checked);
    return 0;
}
