"""Run-to-run reproducibility of C's VALUES on hub tile-rows (R-MAT): step 1 collects their pair lists with atomics
(k_s1_heavy) and then restores ascending-A-tile order -- insertion sort up to 64 entries, csrc/pair_sort.h beyond -- so the
FP64 summation order of every C entry is fixed (the serial SPA's: oracle/spa_ref.c, reference
src/spgemm_serialref_spa_new.h). General positive values, so a different order would show in the last bits."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from spgemm_b200 import api, matrices as M

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


def test_hub_tile_rows_give_bit_identical_values_run_to_run():
    api.init(0)
    m, n, rp, ci, _ = M.rmat(13, 16, seed=3)
    v = M.set_values(len(ci), "hash")
    A = (rp, ci, v)
    B = orc.transpose(m, n, rp, ci, v)
    oA = orc.csr2tile_row_major(m, n, *A)
    assert oA.tile_ptr[1] - oA.tile_ptr[0] > 64, "C tile (0,0) of A A^T must have a pair list the insertion sort does not take"
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    dT = api.transpose(d)
    tA, tB = api.csr2tile(d, False), api.csr2tile(dT, True)
    assert api.tilerow_weights(tA, tB).max() > 2048, "the heavy step-1 path must run"
    runs = []
    for _ in range(3):
        tC, _ = api.spgemm(tA, tB)
        csr = api.tile2csr_device(tC)
        runs.append(csr.download())
        csr.free()
        tC.free()
    for o in (tA, tB, d, dT):
        o.free()
    er, ec, ev = orc.spgemm_spa(A, B, m)
    for r, c, vv in runs:
        assert np.array_equal(r, er) and np.array_equal(c, ec)
        assert np.max(np.abs(vv - ev) / np.abs(ev)) <= 1e-12
    assert np.array_equal(runs[0][2], runs[1][2]) and np.array_equal(runs[0][2], runs[2][2]), "values differ run to run"
