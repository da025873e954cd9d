// tile2csr_emul.cpp -- serial host emulation of tile2csr_v2.cuh (see emul.h). NOT part of the product.
#include <vector>
#include "tile2csr_v2.cuh"

static int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// A slab of C tile-rows [trow0, trow0 + tilem): m rows. Outputs rowptr[m+1] (+ base), out_col / out_val [nnz].
extern "C" int emul_tile2csr(int m, int tilem, int trow0, int numtile, const int *tile_ptr, const int *tile_rowidx, const int *tile_col,
                             const int *tile_nnz, const uint16_t *ptr, const uint16_t *col, const double *val, int base, int *rowptr,
                             int *out_col, double *out_val)
{
    (void)tilem;
    std::vector<int> counts((size_t)16 * numtile + 1, 0), offs((size_t)16 * numtile + 1, 0);
    if (numtile)
        LAUNCH(t2c::k_counts, ceil_div((long long)numtile * 16, 256), 256, numtile, tile_ptr, tile_rowidx, trow0, tile_nnz, ptr, counts.data());
    long long run = 0;
    for (size_t i = 0; i <= (size_t)16 * numtile; i++) { offs[i] = (int)run; run += counts[i]; }   // device: exclusive_scan (n+1 outputs)
    LAUNCH(t2c::k_rowptr, ceil_div((long long)m + 1, 256), 256, m, tile_ptr, offs.data(), numtile, base, rowptr);
    if (numtile)
        LAUNCH(t2c::k_fill, ceil_div((long long)numtile * 16, 256), 256, numtile, tile_ptr, tile_rowidx, trow0, tile_col, tile_nnz, ptr, col,
               val, offs.data(), out_col, out_val);
    return 0;
}
