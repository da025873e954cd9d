// rowplans_emul.cpp -- serial host emulation of rowplans.cuh (see emul.h). NOT part of the product.
#include <vector>
#include "rowplans.cuh"

using namespace rowplans;

static int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Step 1 for A tile-rows [trow0, trow0 + ntr). Outputs: c_tile_ptr[ntr+1]; per C tile (capacity capC) column, tile-row, pair range;
// per pair (capacity capP) A tile and B storage id. info = {B-row recipes, A-row recipes, C tiles, pairs}.
// Returns 0, 1 (fail flag: generic step 1 would run) or 2 (capacity).
extern "C" int emul_rowplans(int tilemA, const int *a_tile_ptr, const int *a_tile_col, int tilemB, const int *b_tile_ptr,
                             const int *b_tile_col, const int *b_rm2csc, int trow0, int ntr, int *c_tile_ptr, int *c_tile_col,
                             int *c_tile_row, int *pair_ptr, int *pair_end, int *pair_a, int *pair_b, long long capC, long long capP,
                             long long *info)
{
    (void)tilemA;
    std::vector<unsigned long long> bkeys(RROW_CAP, 0ull), akeys(RROW_CAP, 0ull);
    std::vector<int> bowner(RROW_CAP, NO_OWNER), aowner(RROW_CAP, NO_OWNER), aflags(RROW_CAP + 1, 0), adense(RROW_CAP + 1, 0),
        rep(RROW_CAP / 2 + 1), brow_id(tilemB > 0 ? tilemB : 1), arow_slot(ntr > 0 ? ntr : 1), arow_recipe(ntr > 0 ? ntr : 1), w(ntr + 1, 0),
        wptr(ntr + 1, 0), ccnt(ntr + 1, 0);
    int nb = 0, na_seen = 0, fail = 0;
    if (tilemB) LAUNCH(k_brow_insert, ceil_div(tilemB, 256), 256, tilemB, b_tile_ptr, b_tile_col, bkeys.data(), bowner.data(), &nb, brow_id.data(), &fail);
    if (fail) return 1;
    if (tilemB) LAUNCH(k_brow_verify, ceil_div(tilemB, 256), 256, tilemB, b_tile_ptr, b_tile_col, brow_id.data(), bowner.data(), &fail);
    if (fail) return 1;
    if (ntr)
        LAUNCH(k_arow_insert, ceil_div(ntr, 256), 256, ntr, trow0, a_tile_ptr, a_tile_col, b_tile_ptr, brow_id.data(), akeys.data(),
               aowner.data(), &na_seen, arow_slot.data(), w.data(), &fail);
    if (fail) return 1;
    LAUNCH(k_flags, ceil_div(RROW_CAP, 256), 256, RROW_CAP, aowner.data(), aflags.data());
    int nrec = 0;
    for (int s = 0; s < RROW_CAP; s++) { adense[s] = nrec; nrec += aflags[s]; }                 // device: exclusive_scan
    LAUNCH(k_reps, ceil_div(RROW_CAP, 256), 256, RROW_CAP, aowner.data(), adense.data(), rep.data());
    if (ntr)
        LAUNCH(k_arow_verify, ceil_div(ntr, 256), 256, ntr, trow0, a_tile_ptr, a_tile_col, brow_id.data(), arow_slot.data(), aowner.data(),
               adense.data(), arow_recipe.data(), &fail);
    if (fail) return 1;

    std::vector<int> rp_numJ(nrec + 1), rp_D((size_t)nrec * MAXW + 1), rp_poff((size_t)nrec * (MAXW + 1) + 1);
    std::vector<unsigned> rp_pair((size_t)nrec * MAXW + 1);
    if (nrec)
        LAUNCH(k_rowplan_build, ceil_div(nrec, 64), 64, nrec, trow0, rep.data(), a_tile_ptr, a_tile_col, b_tile_ptr, b_tile_col,
               rp_numJ.data(), rp_D.data(), rp_poff.data(), rp_pair.data());
    if (ntr) LAUNCH(k_row_counts, ceil_div(ntr, 256), 256, ntr, arow_recipe.data(), rp_numJ.data(), ccnt.data());
    long long nC = 0, nP = 0;
    int maxJ = 1, maxW = 1;
    for (int x = 0; x < ntr; x++) {                                                              // device: two exclusive scans + max
        c_tile_ptr[x] = (int)nC; nC += ccnt[x]; wptr[x] = (int)nP; nP += w[x];
        if (ccnt[x] > maxJ) maxJ = ccnt[x];
        if (w[x] > maxW) maxW = w[x];
    }
    c_tile_ptr[ntr] = (int)nC; wptr[ntr] = (int)nP;
    info[0] = nb; info[1] = nrec; info[2] = nC; info[3] = nP;
    if (nC > capC || nP > capP) return 2;
    if (ntr) {
        LAUNCH(k_expand_tiles, ceil_div((long long)ntr * maxJ, 256), 256, ntr, trow0, maxJ, arow_recipe.data(), rp_numJ.data(), rp_D.data(),
               rp_poff.data(), c_tile_ptr, wptr.data(), c_tile_col, c_tile_row, pair_ptr, pair_end);
        LAUNCH(k_expand_pairs, ceil_div((long long)ntr * maxW, 256), 256, ntr, trow0, maxW, arow_recipe.data(), w.data(), rp_pair.data(),
               wptr.data(), a_tile_ptr, a_tile_col, b_tile_ptr, b_rm2csc, pair_a, pair_b);
    }
    return 0;
}
