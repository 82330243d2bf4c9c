// CPU probe: PriorityMap (cityflow_b200/csrc/priority_map.h) against std::map under the host
// engine's access pattern -- insert on spawn, erase on finish, re-use of erased keys, clear -- plus
// the bounded-growth property (table size follows the LIVE count, not the number of inserts ever).
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include "priority_map.h"

int main() {
    cfb::PriorityMap m;
    std::map<int, int> ref;
    std::mt19937 rnd(123);
    std::vector<int> live;
    long ops = 0;
    for (int round = 0; round < 3; ++round) {
        for (int it = 0; it < 400000; ++it) {
            const unsigned r = rnd() % 100;
            if (r < 50 || live.size() < 100) {  // insert (small key space -> frequent re-use of erased keys)
                const int k = (int) (rnd() % 200000) - 100000;
                if (ref.count(k)) {
                    if (m.get(k) != ref[k] || !m.contains(k)) { printf("FAIL get existing\n"); return 1; }
                } else {
                    if (m.get(k) != -1 || m.contains(k)) { printf("FAIL get missing\n"); return 1; }
                    const int v = (int) (rnd() % 1000000);
                    m.prefetch(k);
                    m.insert(k, v);
                    ref[k] = v;
                    live.push_back(k);
                }
            } else {                             // erase a live key
                const size_t i = rnd() % live.size();
                const int k = live[i];
                live[i] = live.back();
                live.pop_back();
                m.erase(k);
                ref.erase(k);
                if (m.get(k) != -1) { printf("FAIL erased key still found\n"); return 1; }
            }
            ++ops;
            if (m.size() != ref.size()) { printf("FAIL size %zu vs %zu\n", m.size(), ref.size()); return 1; }
        }
        auto s = m.sorted();
        if (s.size() != ref.size()) { printf("FAIL sorted size\n"); return 1; }
        size_t i = 0;
        for (auto &kv : ref) {
            if (s[i].first != kv.first || s[i].second != kv.second) { printf("FAIL sorted order\n"); return 1; }
            ++i;
        }
        if (round == 1) { m.clear(); ref.clear(); live.clear(); if (m.size() != 0 || m.get(5) != -1) { printf("FAIL clear\n"); return 1; } }
    }
    // steady churn with ~1000 live keys: the table must stay small however many inserts went through
    cfb::PriorityMap c;
    std::vector<int> q;
    for (int it = 0; it < 3000000; ++it) {
        const int k = (int) rnd();
        if (c.get(k) >= 0) continue;
        c.insert(k, it & 0xffff);
        q.push_back(k);
        if (q.size() > 1000) { c.erase(q[it % q.size()]); q[it % q.size()] = q.back(); q.pop_back(); }
    }
    if (c.capacity() > (1u << 14)) { printf("FAIL table grew to %zu cells for %zu live keys\n", c.capacity(), c.size()); return 1; }
    printf("OK %ld ops, churn capacity %zu\n", ops, c.capacity());
    return 0;
}
