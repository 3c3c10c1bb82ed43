/*
 * mono_alloc.cpp — a monotonic allocator for the list nodes of ORBextractor::DistributeOctTree (TEST INFRASTRUCTURE).
 *
 * The reference orders equal-size quadtree nodes by HEAP ADDRESS: DistributeOctTree sorts pair<int, ExtractorNode*>
 * (ORBextractor.cc:615, 705-708) and expands them from the back, so both the surviving node set (when the `size >= N`
 * break falls inside a group of equal sizes, :754-755) and the final list ORDER depend on where malloc happened to put
 * each std::list node.  The reference's result is therefore only defined up to the allocator.  This file gives the
 * reference binary in oracle/_ref a deterministic allocator for exactly those nodes: every
 * operator new(sizeof(std::_List_node<ExtractorNode>)) is served from a bump arena with strictly increasing addresses,
 * so "heap address order" becomes "creation order" — the rule the oracle restatement and the CUDA kernel state as pin
 * (ii).  With it, the unmodified reference code and the oracle must agree bit for bit, order included
 * (tests/test_ref_pin.py); with it switched off (rh_set_monotonic_alloc(0): plain malloc) the differences must be
 * confined to the order inside a level and to levels whose cut fell inside a tie group.
 *
 * Linked into liborb_ref.so with -Bsymbolic: only that library's own allocations come here, and the arena is only used while a
 * harness call that runs the extractor is in flight (rh_alloc_scope).
 */
#include <atomic>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <list>
#include <new>

#include <sys/mman.h>

#include "ORBextractor.h"

namespace {

const size_t kNode = sizeof(std::_List_node<ORB_SLAM2::ExtractorNode>);
/* one bump region per thread (the stereo Frame constructor runs two extractors on two std::threads, Frame.cc:79-82):
 * kSlots regions of kSlot bytes inside one virtual reservation; pages are touched lazily */
const size_t kSlot = (size_t)32 << 20;
const int kSlots = 128;

char* g_base = nullptr;
std::atomic<int> g_on{1};
std::atomic<int> g_scope{0};   // > 0 while a harness call that runs ORBextractor::operator() is in flight (rh_alloc_scope)
std::atomic<long> g_served{0};
std::atomic<unsigned char> g_used[kSlots];

struct Slot {
    int id = -1;
    size_t cur = 0;
    long live = 0;
    ~Slot() { if (id >= 0) g_used[id].store(0, std::memory_order_release); }
};
thread_local Slot t_slot;

char* arena() {
    static char* base = [] {
        void* p = mmap(nullptr, kSlot * kSlots, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        return p == MAP_FAILED ? (char*)nullptr : (char*)p;
    }();
    return base;
}

inline bool in_arena(void* p) { return g_base && (char*)p >= g_base && (char*)p < g_base + kSlot * kSlots; }

void* alloc(size_t n) {
    if (n == kNode && g_on.load(std::memory_order_relaxed) && g_scope.load(std::memory_order_relaxed) > 0) {
        if (!g_base) g_base = arena();
        Slot& s = t_slot;
        if (g_base && s.id < 0) {
            for (int i = 0; i < kSlots; ++i) {
                unsigned char expect = 0;
                if (g_used[i].compare_exchange_strong(expect, 1)) { s.id = i; s.cur = 0; s.live = 0; break; }
            }
        }
        if (g_base && s.id >= 0) {
            if (s.cur + kNode <= kSlot) {
                void* p = g_base + (size_t)s.id * kSlot + s.cur;
                s.cur += kNode;
                s.live++;
                g_served.fetch_add(1, std::memory_order_relaxed);
                return p;
            }
            std::fprintf(stderr, "mono_alloc: arena slot exhausted, falling back to malloc (tie order no longer pinned)\n");
        }
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}

void dealloc(void* p) {
    if (!p) return;
    if (in_arena(p)) {
        /* the list nodes of one DistributeOctTree call are created and destroyed by the same thread and die together when
         * its std::list goes out of scope; once none of this thread's nodes is alive its region starts over */
        Slot& s = t_slot;
        if (s.id >= 0 && (size_t)((char*)p - g_base) / kSlot == (size_t)s.id && --s.live == 0) s.cur = 0;
        return;
    }
    std::free(p);
}

}  // namespace

void* operator new(size_t n) { return alloc(n); }
void* operator new[](size_t n) { return alloc(n); }
void operator delete(void* p) noexcept { dealloc(p); }
void operator delete[](void* p) noexcept { dealloc(p); }
void operator delete(void* p, size_t) noexcept { dealloc(p); }
void operator delete[](void* p, size_t) noexcept { dealloc(p); }

extern "C" {
/* The arena only serves allocations made while an extraction is running: a same-sized allocation elsewhere (a std::string
 * buffer, say) may be released inside libstdc++.so, whose operator delete knows nothing of the arena. */
void rh_alloc_scope(int delta) { g_scope.fetch_add(delta); }
void rh_set_monotonic_alloc(int on) { g_on.store(on ? 1 : 0); }
long rh_monotonic_alloc_served(void) { return g_served.load(); }
}
