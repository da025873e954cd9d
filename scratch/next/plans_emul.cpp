// plans_emul.cpp -- host orchestration of the plans.cuh kernels under serial emulation (emul.h). Mirrors what
// spgemm_device would do between step 1 (pair lists) and the end of step 3. Built by scratch/next/Makefile into
// libplans_emul.so and driven by test_plans_emul.py. NOT part of the product.
#include <stdlib.h>
#include <vector>
#include "plans.cuh"

using namespace plans;

static int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Inputs: A tiles (row-major) and B tiles (CSC storage order) as mask / Ptr / tile_nnz / val; the pair lists of the
// numblkC C tiles (pair_b holds B storage ids). Outputs: C's mask / Ptr [numblkC*16], tile_nnz [numblkC+1] (offsets),
// Col / Val (capacity nnz_cap), info[0..3] = patterns, recipes, plan entries, nnzC. Returns 0, 1 (fell back: fail
// flag raised) or 2 (nnz_cap too small).
extern "C" int emul_plans(int numtileA, const uint16_t *a_mask, const uint16_t *a_ptr, const int *a_tile_nnz, const double *a_val,
                          int numtileB, const uint16_t *b_mask, const uint16_t *b_ptr, const int *b_tile_nnz, const double *b_val,
                          int numblkC, const int *pair_ptr, const int *pair_end, const int *pair_a, const int *pair_b,
                          uint16_t *c_mask, uint16_t *c_ptr, int *c_tile_nnz, uint16_t *c_col, double *c_val, long long nnz_cap,
                          long long *info)
{
    std::vector<unsigned> pstate(PCAP, 0u), rstate(RCAP, 0u);
    std::vector<uint4> pblk(2 * PCAP);
    std::vector<unsigned long long> rhash(RCAP);
    std::vector<int> rrep(RCAP), rdense(RCAP), rep_tile(RCAP / 2), patA(numtileA > 0 ? numtileA : 1), patB(numtileB > 0 ? numtileB : 1),
        recipe_id(numblkC > 0 ? numblkC : 1);
    int npat = 0, nrec = 0, fail = 0;

    if (numtileA) LAUNCH(k_pattern_ids, ceil_div(numtileA, 256), 256, numtileA, a_mask, pstate.data(), pblk.data(), &npat, patA.data(), &fail);
    if (numtileB) LAUNCH(k_pattern_ids, ceil_div(numtileB, 256), 256, numtileB, b_mask, pstate.data(), pblk.data(), &npat, patB.data(), &fail);
    if (fail) return 1;
    if (numblkC)
        LAUNCH(k_recipe_ids, ceil_div(numblkC, 256), 256, numblkC, pair_ptr, pair_end, pair_a, pair_b, patA.data(), patB.data(),
               rstate.data(), rhash.data(), rrep.data(), rdense.data(), &nrec, recipe_id.data(), rep_tile.data(), &fail);
    if (fail) return 1;

    std::vector<uint16_t> plan_mask((size_t)nrec * TS + 1), plan_ptr((size_t)nrec * TS + 1);
    std::vector<int> plan_nnz(nrec + 1), plan_tot(nrec + 1), plan_off(nrec + 1);
    std::vector<unsigned> plan_start((size_t)nrec * PLAN_ROWS + 1);
    std::vector<uint8_t> plan_col((size_t)nrec * 256 + 1);
    if (nrec)
        LAUNCH(k_plan_build<false>, ceil_div(nrec, 64), 64, nrec, rep_tile.data(), pair_ptr, pair_end, pair_a, pair_b, a_mask, a_ptr, b_mask,
               b_ptr, plan_mask.data(), plan_ptr.data(), plan_nnz.data(), plan_tot.data(), (const int *)nullptr, (unsigned *)nullptr,
               (uint8_t *)nullptr, (unsigned *)nullptr);
    long long entries = 0;
    for (int r = 0; r < nrec; r++) { plan_off[r] = (int)entries; entries += plan_tot[r]; }   // device: exclusive_scan
    std::vector<unsigned> plan_ent((size_t)entries + 1);
    if (nrec)
        LAUNCH(k_plan_build<true>, ceil_div(nrec, 64), 64, nrec, rep_tile.data(), pair_ptr, pair_end, pair_a, pair_b, a_mask, a_ptr, b_mask,
               b_ptr, plan_mask.data(), plan_ptr.data(), plan_nnz.data(), plan_tot.data(), plan_off.data(), plan_start.data(),
               plan_col.data(), plan_ent.data());

    std::vector<int> cnt(numblkC + 1, 0);
    if (numblkC)
        LAUNCH(k_symbolic_from_plans, ceil_div((long long)numblkC * 16, 256), 256, numblkC, recipe_id.data(), plan_mask.data(),
               plan_ptr.data(), plan_nnz.data(), c_mask, c_ptr, cnt.data());
    long long nnzC = 0;
    for (int t = 0; t < numblkC; t++) { c_tile_nnz[t] = (int)nnzC; nnzC += cnt[t]; }         // device: exclusive_scan
    c_tile_nnz[numblkC] = (int)nnzC;
    info[0] = npat; info[1] = nrec; info[2] = entries; info[3] = nnzC;
    if (nnzC > nnz_cap) return 2;

    std::vector<int> blk2tile((size_t)((nnzC + 31) >> 5) + 1, 0);
    for (int t = 0; t < numblkC; t++)                                                           // device: k_blk2tile
        for (long long blk = ((long long)c_tile_nnz[t] + 31) >> 5; (blk << 5) < c_tile_nnz[t + 1]; blk++) blk2tile[blk] = t;
    if (nnzC)
        LAUNCH(k_numeric_from_plans, ceil_div(nnzC, 256), 256, numblkC, (int)nnzC, blk2tile.data(), c_tile_nnz, recipe_id.data(),
               plan_start.data(), plan_col.data(), plan_ent.data(), pair_ptr, pair_a, pair_b, a_tile_nnz, a_val, b_tile_nnz, b_val, c_col,
               c_val);
    return 0;
}
