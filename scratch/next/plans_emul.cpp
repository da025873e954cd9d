// plans_emul.cpp -- host orchestration of the plans.cuh kernels under serial emulation (emul.h). Mirrors what
// spgemm_device would do between step 1 (pair lists) and the end of step 3. Built by scratch/next/Makefile into
// libplans_emul.so and driven by test_plans_emul.py. NOT part of the product.
#include <stdlib.h>
#include <string.h>
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
    std::vector<unsigned long long> pkeys(PCAP, 0ull), rkeys(RCAP, 0ull);
    std::vector<int> powner(PCAP, NO_OWNER), rowner(RCAP, NO_OWNER), rflags(RCAP + 1, 0), rdense(RCAP + 1, 0), rep_tile(RCAP / 2 + 1),
        patA(numtileA > 0 ? numtileA : 1), patB(numtileB > 0 ? numtileB : 1), rslot(numblkC > 0 ? numblkC : 1), recipe_id(numblkC > 0 ? numblkC : 1);
    int npat = 0, nrec_seen = 0, fail = 0;

    if (numtileA) LAUNCH(k_pattern_insert, ceil_div(numtileA, 256), 256, numtileA, 0, a_mask, pkeys.data(), powner.data(), &npat, patA.data(), &fail);
    if (numtileB) LAUNCH(k_pattern_insert, ceil_div(numtileB, 256), 256, numtileB, numtileA, b_mask, pkeys.data(), powner.data(), &npat, patB.data(), &fail);
    if (fail) return 1;
    if (numtileA) LAUNCH(k_pattern_verify, ceil_div(numtileA, 256), 256, numtileA, a_mask, patA.data(), powner.data(), numtileA, a_mask, b_mask, &fail);
    if (numtileB) LAUNCH(k_pattern_verify, ceil_div(numtileB, 256), 256, numtileB, b_mask, patB.data(), powner.data(), numtileA, a_mask, b_mask, &fail);
    if (fail) return 1;
    if (numblkC)
        LAUNCH(k_recipe_insert, ceil_div(numblkC, 256), 256, numblkC, pair_ptr, pair_end, pair_a, pair_b, patA.data(), patB.data(),
               rkeys.data(), rowner.data(), &nrec_seen, rslot.data(), &fail);
    if (fail) return 1;
    LAUNCH(k_recipe_flags, ceil_div(RCAP, 256), 256, rowner.data(), rflags.data());
    int nrec = 0;
    for (int s = 0; s < RCAP; s++) { rdense[s] = nrec; nrec += rflags[s]; }                   // device: exclusive_scan
    LAUNCH(k_recipe_reps, ceil_div(RCAP, 256), 256, rowner.data(), rdense.data(), rep_tile.data());
    if (numblkC)
        LAUNCH(k_recipe_verify, ceil_div(numblkC, 256), 256, numblkC, pair_ptr, pair_end, pair_a, pair_b, patA.data(), patB.data(),
               rslot.data(), rowner.data(), rdense.data(), recipe_id.data(), &fail);
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
    // variant B (per-pair value bases) must give the same bits
    const long long npairs = numblkC ? pair_end[numblkC - 1] : 0;
    std::vector<int2_> pair_base((size_t)npairs + 1);
    std::vector<uint16_t> col_b((size_t)nnzC + 1);
    std::vector<double> val_b((size_t)nnzC + 1);
    if (npairs) LAUNCH(k_pair_bases, ceil_div(npairs, 256), 256, npairs, pair_a, pair_b, a_tile_nnz, b_tile_nnz, pair_base.data());
    if (nnzC)
        LAUNCH(k_numeric_from_plans_b, ceil_div(nnzC, 256), 256, numblkC, (int)nnzC, blk2tile.data(), c_tile_nnz, recipe_id.data(),
               plan_start.data(), plan_col.data(), plan_ent.data(), pair_ptr, pair_base.data(), a_val, b_val, col_b.data(), val_b.data());
    for (long long g = 0; g < nnzC; g++)
        if (col_b[g] != c_col[g] || memcmp(&val_b[g], &c_val[g], sizeof(double))) return 3;
    return 0;
}
