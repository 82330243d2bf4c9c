// priority_map.h -- the host engine's vehiclePool index (Engine::vehiclePool, engine.h:25).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <new>
#include <utility>
#include <vector>
#include <sys/mman.h>

namespace cfb {

// Allocator for the two tables the spawn loop hits at random (priority table, slot records: tens of MB at 1e6 vehicles):
// 2 MiB-aligned blocks with MADV_HUGEPAGE, so that a random access costs a cache miss but not a TLB miss on top.
template <class T>
struct HugePageAllocator {
    using value_type = T;
    HugePageAllocator() = default;
    template <class U> HugePageAllocator(const HugePageAllocator<U> &) {}
    T *allocate(size_t n) {
        const size_t bytes = n * sizeof(T);
        if (bytes < ((size_t) 1 << 21)) return static_cast<T *>(::operator new(bytes));
        const size_t rounded = (bytes + ((size_t) 1 << 21) - 1) & ~(((size_t) 1 << 21) - 1);
        void *p = std::aligned_alloc((size_t) 1 << 21, rounded);
        if (!p) throw std::bad_alloc();
        madvise(p, rounded, MADV_HUGEPAGE);
        return static_cast<T *>(p);
    }
    void deallocate(T *p, size_t n) {
        if (n * sizeof(T) < ((size_t) 1 << 21)) ::operator delete(p); else std::free(p);
    }
    template <class U> bool operator==(const HugePageAllocator<U> &) const { return true; }
    template <class U> bool operator!=(const HugePageAllocator<U> &) const { return false; }
};

// Open-addressing hash map priority -> slot: Engine::checkPriority (engine.cpp:601) is on the
// per-spawn path; the reference's ordered std::map is only materialised (sorted) when an API call
// needs vehiclePool order.
class PriorityMap {
public:
    PriorityMap() { rehash(1 << 12); }
    bool contains(int k) const { return find(k) >= 0; }
    int get(int k) const { long i = find(k); return i >= 0 ? cell_[i].val : -1; }
    void insert(int k, int v) {
        if ((size_ + 1) * 4 > cap_) rehash(roomFor(size_ + 1));
        size_t i = hash(k);
        while (cell_[i].val >= 0) i = (i + 1) & (cap_ - 1);
        cell_[i].key = k; cell_[i].val = v; ++size_;
    }
    // Backward-shift deletion: the entries behind the hole move up while that keeps them reachable from their home
    // cell, so there are no tombstones and a table whose population is steady (vehicles arrive as fast as they leave)
    // never has to be rebuilt.
    void erase(int k) {
        long f = find(k);
        if (f < 0) return;
        size_t i = (size_t) f, j = i;
        const size_t mask = cap_ - 1;
        for (;;) {
            j = (j + 1) & mask;
            if (cell_[j].val < 0) break;
            const size_t h = hash(cell_[j].key);
            // cell j may move into the hole i unless its home h lies cyclically in (i, j]
            const bool homeBetween = i <= j ? (h > i && h <= j) : (h > i || h <= j);
            if (!homeBetween) { cell_[i] = cell_[j]; i = j; }
        }
        cell_[i].val = EMPTY;
        --size_;
    }
    void clear() { rehash(1 << 12, false); }
    size_t size() const { return size_; }
    size_t capacity() const { return cap_; }
    // the spawn loop knows the keys it is about to look up (the RNG can be run ahead): start the
    // one cache miss a lookup costs early
    void prefetch(int k) const { __builtin_prefetch(&cell_[hash(k)]); }
    // (priority, slot) pairs in ascending priority = vehiclePool iteration order
    std::vector<std::pair<int, int>> sorted() const {
        std::vector<std::pair<int, int>> out;
        out.reserve(size_);
        for (size_t i = 0; i < cap_; ++i) if (cell_[i].val >= 0) out.emplace_back(cell_[i].key, cell_[i].val);
        std::sort(out.begin(), out.end());
        return out;
    }
private:
    enum { EMPTY = -1 };
    struct Cell { int key, val; };   // val >= 0: slot; one 8-byte cell = one cache line touched per probe
    size_t hash(int k) const { return ((uint32_t) k * 2654435761u) & (cap_ - 1); }
    long find(int k) const {
        size_t i = hash(k);
        while (cell_[i].val != EMPTY) {
            if (cell_[i].key == k) return (long) i;
            i = (i + 1) & (cap_ - 1);
        }
        return -1;
    }
    // table size for n live entries: load <= 1/4 (grown only when the population grows)
    static size_t roomFor(size_t n) {
        size_t c = 1 << 12;
        while (c < n * 4) c *= 2;
        return c;
    }
    void rehash(size_t n, bool keep = true) {
        std::vector<Cell, HugePageAllocator<Cell>> old = std::move(cell_);
        cap_ = n; cell_.assign(n, Cell{0, EMPTY}); size_ = 0;
        if (keep) for (const Cell &c : old) if (c.val >= 0) insert(c.key, c.val);
    }
    std::vector<Cell, HugePageAllocator<Cell>> cell_;
    size_t cap_ = 0, size_ = 0;
};

// The same map cut into four by the top two bits of the (uniformly random) priority, so that four threads can each own a
// quarter of it while a step's vehicles are created (host_engine.cpp, parallel creation); everywhere else it behaves
// like one PriorityMap.
class PriorityMap4 {
public:
    static int part(int k) { return (int) ((uint32_t) k >> 30); }
    PriorityMap &sub(int p) { return m_[p].m; }
    bool contains(int k) const { return m_[part(k)].m.contains(k); }
    int get(int k) const { return m_[part(k)].m.get(k); }
    void insert(int k, int v) { m_[part(k)].m.insert(k, v); }
    void erase(int k) { m_[part(k)].m.erase(k); }
    void clear() { for (auto &m : m_) m.m.clear(); }
    size_t size() const { size_t n = 0; for (auto &m : m_) n += m.m.size(); return n; }
    size_t capacity() const { size_t n = 0; for (auto &m : m_) n += m.m.capacity(); return n; }
    void prefetch(int k) const { m_[part(k)].m.prefetch(k); }
    std::vector<std::pair<int, int>> sorted() const {
        std::vector<std::pair<int, int>> out;
        for (auto &m : m_) { auto s = m.m.sorted(); out.insert(out.end(), s.begin(), s.end()); }
        std::sort(out.begin(), out.end());
        return out;
    }
private:
    struct alignas(128) Padded { PriorityMap m; };   // each quarter's size counter on its own cache line (four threads insert at once)
    Padded m_[4];
};

}  // namespace cfb
