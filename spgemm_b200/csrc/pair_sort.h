// pair_sort.h -- in-place heap sort of a pair list by its A-tile index (the keys of one list are distinct: an A tile
// meets at most one B tile per C tile). One thread sorts one list; no recursion, no extra storage. k_s1_heavy uses it
// for the lists its insertion sort does not take, so that the FP64 summation order of every C entry is the serial
// SPA's (ascending A tile) and values are reproducible run to run. Compiles as plain C++ too: tests/test_pair_sort.py
// runs exactly this code on the CPU.
#pragma once
#if defined(__CUDACC__)
#define TSG_HD __host__ __device__ __forceinline__
#else
#define TSG_HD static inline
#endif

namespace tsg {

TSG_HD void pair_sift_down(int *ka, int *kb, int root, int n)
{
    const int a = ka[root], b = kb[root];
    for (;;) {
        int child = 2 * root + 1;
        if (child >= n) break;
        int ca = ka[child];
        if (child + 1 < n) {
            const int ca2 = ka[child + 1];
            if (ca2 > ca) { child++; ca = ca2; }
        }
        if (ca <= a) break;
        ka[root] = ca; kb[root] = kb[child];
        root = child;
    }
    ka[root] = a; kb[root] = b;
}

// ascending by ka; kb follows
TSG_HD void pair_heap_sort(int *ka, int *kb, int n)
{
    if (n < 2) return;
    for (int r = n / 2 - 1; r >= 0; r--) pair_sift_down(ka, kb, r, n);
    for (int end = n - 1; end > 0; end--) {
        const int a = ka[0], b = kb[0];
        ka[0] = ka[end]; kb[0] = kb[end];
        ka[end] = a; kb[end] = b;
        pair_sift_down(ka, kb, 0, end);
    }
}

}  // namespace tsg
