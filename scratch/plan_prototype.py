"""CPU prototype of the round-2 "recipe plan" symbolic + numeric steps, checked against the oracle.

A C tile's recipe = the sequence of (A-tile pattern, B-tile pattern) over its pairs in ascending K. For every distinct recipe one
plan is built from a representative tile: the C tile's row masks / Ptr / nnz and, per C nonzero in storage order, the sources
(pair index, position in A's tile, position in B's tile) in the serial SPA's order (ascending pair, then ascending k).
Then: symbolic = copy the plan's masks; numeric = walk the plan. Compared bit-exactly (integer-valued inputs) with the oracle's C.

usage: python scratch/plan_prototype.py stencil27:12 | lap2d:40 | blockfem:200 | stencil27x:20,7,5
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.getcwd())
from oracle import oracle as orc  # noqa: E402
from spgemm_b200 import matrices as M  # noqa: E402

kind, arg = sys.argv[1].split(":")
if kind == "stencil27x":
    m, n, rp, ci, _ = M.stencil27(*[int(x) for x in arg.split(",")])
else:
    m, n, rp, ci, _ = {"stencil27": M.stencil27, "lap2d": M.lap2d, "blockfem": M.blockfem}[kind](int(arg))
v = M.set_values(len(ci), "mod10")
A = (rp, ci, v)
t0 = time.time()
tA = orc.csr2tile_row_major(m, n, *A)            # B = A; the per-tile CSR of a tile is the same in both storage orders
csrC = orc.spgemm_spa(A, A, n)
tC = orc.ctiles_from_csr(m, n, tA, orc.csr2tile_col_major(m, n, *A), csrC)
nt = tA.numtile
mask = tA.mask.reshape(nt, 16).astype(np.uint32)
ptr = tA.ptr.reshape(nt, 16).astype(np.int64)
base = tA.tile_nnz.astype(np.int64)

# ---- pattern ids (device: hash of the 32-byte mask block + small hash table)
_, pat = np.unique(np.ascontiguousarray(tA.mask.reshape(nt, 16)).view(np.dtype((np.void, 32))).ravel(), return_inverse=True)

# ---- pair lists per C tile, ascending K (device: step 1 already emits exactly this)
tile_ptr, tile_col = tA.tile_ptr.astype(np.int64), tA.tile_columnidx.astype(np.int64)
tile_row = np.repeat(np.arange(tA.tilem, dtype=np.int64), np.diff(tile_ptr))
cnt = tile_ptr[tile_col + 1] - tile_ptr[tile_col]
src = np.repeat(np.arange(nt, dtype=np.int64), cnt)
dst = np.repeat(tile_ptr[tile_col], cnt) + (np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt))
ckey = tile_row[src] * tA.tilen + tile_col[dst]
order = np.lexsort((tile_col[src], ckey))
ckey, pair_a, pair_b = ckey[order], src[order], dst[order]
first = np.flatnonzero(np.r_[True, ckey[1:] != ckey[:-1]])
pair_ptr = np.r_[first, len(ckey)]
nC = len(first)
assert nC == tC.numtile, (nC, tC.numtile)
# the oracle lists C tiles in (tile-row, tile-column) order = ascending ckey: same order as ours
assert np.array_equal(ckey[first] // tA.tilen, tC.tile_rowidx) and np.array_equal(ckey[first] % tA.tilen, tC.tile_columnidx)

# ---- recipe ids (device: hash of the (patA, patB) sequence + small hash table)
recipe_of = {}
recipe_id = np.zeros(nC, np.int64)
rep = []                                            # representative C tile of each recipe
for t in range(nC):
    key = tuple(zip(pat[pair_a[pair_ptr[t]:pair_ptr[t + 1]]].tolist(), pat[pair_b[pair_ptr[t]:pair_ptr[t + 1]]].tolist()))
    if key not in recipe_of:
        recipe_of[key] = len(rep)
        rep.append(t)
    recipe_id[t] = recipe_of[key]
print(f"{sys.argv[1]}: A tiles {nt}, patterns {pat.max() + 1}, C tiles {nC}, recipes {len(rep)}")


def bits(mk):                                       # columns of a 16-bit row mask, ascending (bit 15-c <-> column c)
    return [c for c in range(16) if mk & (0x8000 >> c)]


# ---- plans: one per recipe, from its representative tile
plans = []
for t in rep:
    pa, pb = pair_a[pair_ptr[t]:pair_ptr[t + 1]], pair_b[pair_ptr[t]:pair_ptr[t + 1]]
    cm = np.zeros(16, np.uint32)
    for a, b in zip(pa, pb):
        for r in range(16):
            for k in bits(mask[a, r]):
                cm[r] |= mask[b, k]
    cptr = np.zeros(16, np.int64)
    run = 0
    srcs = []                                       # per C nonzero (storage order): [(pair, posA, posB), ...]
    cols = []
    for r in range(16):
        cptr[r] = run
        for c in bits(cm[r]):
            lst = []
            for p, (a, b) in enumerate(zip(pa, pb)):
                for ia, k in enumerate(bits(mask[a, r])):
                    if mask[b, k] & (0x8000 >> c):
                        posb = ptr[b, k] + bin(int(mask[b, k]) >> (16 - c)).count("1")
                        lst.append((p, int(ptr[a, r]) + ia, int(posb)))
            srcs.append(lst)
            cols.append(c)
            run += 1
    plans.append((cm, cptr, run, cols, srcs))

# ---- symbolic from plans == oracle's symbolic
c_mask = np.stack([plans[r][0] for r in recipe_id]) if nC else np.zeros((0, 16), np.uint32)
c_ptr = np.stack([plans[r][1] for r in recipe_id]) if nC else np.zeros((0, 16), np.int64)
c_nnz = np.array([plans[r][2] for r in recipe_id], np.int64)
assert np.array_equal(c_mask.astype(np.uint16).ravel(), tC.mask), "mask"
assert np.array_equal(c_ptr.astype(np.uint16).ravel(), tC.ptr), "Ptr"
assert np.array_equal(np.r_[0, np.cumsum(c_nnz)], tC.tile_nnz), "tile_nnz"

# ---- numeric from plans == oracle's values (vectorised per recipe over all its tiles)
c_base = np.r_[0, np.cumsum(c_nnz)]
val = np.zeros(int(c_base[-1]))
col = np.zeros(int(c_base[-1]), np.uint16)
nsrc = 0
for rid, (cm, cptr, run, cols, srcs) in enumerate(plans):
    tiles = np.flatnonzero(recipe_id == rid)
    pp = pair_ptr[tiles]
    for j in range(run):
        acc = np.zeros(len(tiles))
        for p, posa, posb in srcs[j]:
            a, b = pair_a[pp + p], pair_b[pp + p]
            acc = acc + tA.val[base[a] + posa] * tA.val[base[b] + posb]   # integer-valued: exact in any order
            nsrc += len(tiles)
        val[c_base[tiles] + j] = acc
        col[c_base[tiles] + j] = cols[j]
assert np.array_equal(col, tC.col), "Col"
assert np.array_equal(val, tC.val), "values"
plan_entries = sum(len(s) for pl in plans for s in pl[4])
print(f"  plans reproduce the oracle's C bit-exactly: {int(c_base[-1])} nonzeros, {nsrc} products (nnzCub {orc.nnzcub(ci, rp)}), "
      f"{plan_entries} plan entries in total ({plan_entries * 4 / 1024:.1f} KB at 4 B each), {time.time() - t0:.1f}s")
