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
 * Linked into liborb_ref.so with -Bsymbolic: only that library's own allocations come here.
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
const size_t kArena = (size_t)256 << 20;   /* virtual reservation; pages are touched lazily */

char* g_base = nullptr;
std::atomic<size_t> g_cur{0};
std::atomic<long> g_live{0};
std::atomic<int> g_on{1};
std::atomic<long> g_served{0};

char* arena() {
    static char* base = [] {
        void* p = mmap(nullptr, kArena, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        return p == MAP_FAILED ? (char*)nullptr : (char*)p;
    }();
    return base;
}

inline bool in_arena(void* p) { return g_base && (char*)p >= g_base && (char*)p < g_base + kArena; }

void* alloc(size_t n) {
    if (n == kNode && g_on.load(std::memory_order_relaxed)) {
        if (!g_base) g_base = arena();
        if (g_base) {
            const size_t off = g_cur.fetch_add(kNode, std::memory_order_relaxed);
            if (off + kNode <= kArena) {
                g_live.fetch_add(1, std::memory_order_relaxed);
                g_served.fetch_add(1, std::memory_order_relaxed);
                return g_base + off;
            }
            std::fprintf(stderr, "mono_alloc: arena exhausted, falling back to malloc (tie order no longer pinned)\n");
        }
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}

void dealloc(void* p) {
    if (!p) return;
    if (in_arena(p)) {
        /* all list nodes of a DistributeOctTree call die together when its std::list goes out of scope; once none is
         * alive (in any thread) the arena starts over */
        if (g_live.fetch_sub(1, std::memory_order_acq_rel) == 1) g_cur.store(0, std::memory_order_release);
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
void rh_set_monotonic_alloc(int on) { g_on.store(on ? 1 : 0); }
long rh_monotonic_alloc_served(void) { return g_served.load(); }
}
