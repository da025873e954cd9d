"""How repetitive is the tile structure of the BASELINE configs?  (CPU only; oracle tiles.)
Counts distinct A-tile patterns (the 16 row masks) and distinct C-tile "recipes" (the sequence of
(A pattern, B pattern) over the C tile's pairs in ascending K) -- the quantity a pattern-cached numeric
step would key its plans on.   usage: python scratch/pattern_stats.py stencil27:64 | lap2d:256 | blockfem:20000 | rmat:14"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.getcwd())
from oracle import oracle as orc
from spgemm_b200 import matrices as M

kind, arg = sys.argv[1].split(":")
arg = int(arg)
gen = {"stencil27": lambda: M.stencil27(arg), "lap2d": lambda: M.lap2d(arg), "blockfem": lambda: M.blockfem(arg),
       "rmat": lambda: M.rmat(arg, 16, seed=1), "rmatmild": lambda: M.rmat(arg, 16, a=.3, b=.25, c=.25, d=.2, seed=1)}[kind]
m, n, rp, ci, v = gen()
t0 = time.time()
tA = orc.csr2tile_row_major(m, n, rp, ci, v)
nt = tA.numtile
masks = np.ascontiguousarray(tA.mask.reshape(nt, 16))
_, pat = np.unique(masks.view(np.dtype((np.void, 32))).ravel(), return_inverse=True)
npat = int(pat.max()) + 1
print(f"{sys.argv[1]}: m={m} nnz={len(ci)} A tiles={nt} distinct A-tile patterns={npat} ({time.time()-t0:.1f}s)")

# pairs of C = A*A: for A tile t=(I,K), every A tile u=(K,J) of tile-row K. B tile pattern = pattern of u (same matrix).
tile_ptr = tA.tile_ptr.astype(np.int64)
tile_col = tA.tile_columnidx.astype(np.int64)
tile_row = np.repeat(np.arange(tA.tilem, dtype=np.int64), np.diff(tile_ptr))
cnt = (tile_ptr[tile_col + 1] - tile_ptr[tile_col])            # pairs started by each A tile
npairs = int(cnt.sum())
src = np.repeat(np.arange(nt, dtype=np.int64), cnt)               # A tile of each pair
start = np.repeat(tile_ptr[tile_col], cnt)
within = np.arange(npairs, dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
dst = start + within                                              # B tile of each pair
I, J, K = tile_row[src], tile_col[dst], tile_col[src]
ckey = I * tA.tilen + J
order = np.lexsort((K, ckey))                                     # group by C tile, ascending K inside
ckey, pa, pb = ckey[order], pat[src[order]].astype(np.uint64), pat[dst[order]].astype(np.uint64)
first = np.r_[True, ckey[1:] != ckey[:-1]]
gid = np.cumsum(first) - 1
nC = int(gid[-1]) + 1
pos = np.arange(npairs) - np.flatnonzero(first)[gid]              # index of the pair inside its C tile
with np.errstate(over="ignore"):
    item = (pa * np.uint64(npat) + pb + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    w = np.power(np.uint64(0xD6E8FEB86659FD93), (pos + 1).astype(np.uint64))
    h = np.add.reduceat(item * w, np.flatnonzero(first))
    h ^= np.add.reduceat(np.ones(npairs, np.uint64), np.flatnonzero(first)) << np.uint64(56)
uniq, inv, counts = np.unique(h, return_inverse=True, return_counts=True)
plen = np.add.reduceat(np.ones(npairs, np.int64), np.flatnonzero(first))
print(f"  tile pairs={npairs} listed C tiles={nC} pairs/C tile={npairs/nC:.2f} distinct C-tile recipes={len(uniq)} "
      f"({nC/len(uniq):.1f} C tiles per recipe; top recipe covers {counts.max()/nC:.1%}, top 100 cover {np.sort(counts)[-100:].sum()/nC:.1%})")
# pairs in recipes = plan size if every distinct recipe is planned once
first_of_recipe = np.unique(inv, return_index=True)[1]
print(f"  pairs to plan once per recipe={int(plen[first_of_recipe].sum())} ({plen[first_of_recipe].sum()/npairs:.2%} of all pairs)  total {time.time()-t0:.1f}s")
