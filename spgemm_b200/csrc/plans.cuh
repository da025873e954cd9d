// plans.cuh -- constants and the wait-free hash-table insert shared by plans.cu and the recipe hashing inside k_s1_fill.
#pragma once
#include "common.cuh"

namespace tsg {
namespace plans {

constexpr int PCAP = 1 << 13;            // pattern table slots; more than PCAP/2 distinct patterns => no pattern ids
constexpr int PLANS_MAX_PATTERNS = 1024; // the plan path is attempted only when both operands hold at most this many patterns
constexpr int RCAP = 1 << 14;            // recipe table slots
constexpr int RMAX = RCAP / 2;           // more distinct recipes than this => fail (generic kernels run)
constexpr int PLAN_ENT_CAP = 1 << 24;    // plan entries (products of all distinct recipes) the buffer holds
constexpr int PLANS_CHAIN_MIN = 12;      // k_plan_slots: chains of C nonzeros are built to at least this many products
constexpr int NO_OWNER = 0x7f7f7f7f;     // what cudaMemset(0x7f) leaves; item indices stay below it

__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v)
{
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    return h;
}

// Deduplication without waiting: insert the 64-bit hash with one atomicCAS on the key word; the smallest item index
// that lands in a slot becomes its owner (atomicMin, done by the caller: deterministic); every item later compares its
// full key with the owner's. Returns the slot, or -1 after raising *fail (table full / too many distinct keys).
// Millions of items, a few dozen hot slots: the key is read with a plain (L1-cached) load first. A slot only ever goes
// from empty to one key, so a stale read can only show "empty" -- and then the atomicCAS returns the truth.
__device__ __forceinline__ int table_insert(unsigned long long *keys, int cap, unsigned long long h, int *count, int limit, int *fail)
{
    if (h == 0ull) h = 1ull;  // 0 marks an empty slot
    unsigned slot = (unsigned)(h >> 20) & (unsigned)(cap - 1);
    for (int probe = 0; probe < cap; probe++, slot = (slot + 1) & (unsigned)(cap - 1)) {
        if (probe && *(volatile int *)fail) return -1;  // a table that overflowed is not worth probing to the end
        unsigned long long old = keys[slot];
        if (old == 0ull) old = atomicCAS(&keys[slot], 0ull, h);
        if (old == 0ull) {
            if (atomicAdd(count, 1) >= limit) *fail = 1;
            return (int)slot;
        }
        if (old == h) return (int)slot;
    }
    *fail = 1;
    return -1;
}

}  // namespace plans

// the recipe table as k_s1_fill sees it
struct PlanTable {
    unsigned long long *keys;
    int *owner, *count, *fail;
};

}  // namespace tsg
