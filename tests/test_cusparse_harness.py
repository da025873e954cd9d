"""cuSPARSE as a second, independent opinion on C = A*B (the reference's own check, src/external/cusparse/spgemm_cusparse.h,
with its CUDA_R_32F descriptor bug fixed). Test-side only: cuSPARSE is never on the product path."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT
from spgemm_b200 import api, matrices as M

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available() or shutil.which("nvcc") is None, reason="needs a GPU and nvcc")]


def write_csr(path, m, n, rp, ci, v):
    with open(path, "wb") as f:
        np.array([m, n, len(ci)], np.int32).tofile(f)
        np.asarray(rp, np.int32).tofile(f)
        np.asarray(ci, np.int32).tofile(f)
        np.asarray(v, np.float64).tofile(f)


def read_csr(path):
    with open(path, "rb") as f:
        m, n, nnz = np.fromfile(f, np.int32, 3)
        return np.fromfile(f, np.int32, m + 1), np.fromfile(f, np.int32, nnz), np.fromfile(f, np.float64, nnz)


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("cusparse") / "cusparse_check"
    subprocess.check_call(["nvcc", "-O2", "-o", str(out), os.path.join(ROOT, "tests", "cusparse", "cusparse_check.cu"), "-lcusparse"])
    return str(out)


@pytest.mark.parametrize("name", ["stencil27_14", "rmat_s11_aat", "rect"])
def test_against_cusparse(exe, tmp_path, name):
    api.init(0)
    if name == "stencil27_14":
        m, k, rp, ci, _ = M.stencil27(14)
        A = B = (rp, ci, M.set_values(len(ci), "hash"))
        n = k
    elif name == "rmat_s11_aat":
        m, k, rp, ci, _ = M.rmat(11, 8, seed=4)
        A = (rp, ci, M.set_values(len(ci), "hash"))
        d = api.DeviceCSR.upload(m, k, *A)
        dT = api.transpose(d)
        B = dT.download()
        n = m
        dT.free(); d.free()
    else:
        m, k, rpA, ciA, vA = M.random_sparse(300, 210, 0.03, seed=5)
        _, n, rpB, ciB, vB = M.random_sparse(210, 170, 0.04, seed=6)
        A, B = (rpA, ciA, vA), (rpB, ciB, vB)
    write_csr(tmp_path / "A.csr", m, k, *A)
    write_csr(tmp_path / "B.csr", k, n, *B)
    out = subprocess.run([exe, str(tmp_path / "A.csr"), str(tmp_path / "B.csr"), str(tmp_path / "C.csr")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    crp, cci, cv = read_csr(tmp_path / "C.csr")
    r, c, v, st = api.spgemm_csr_host(m, k, n, A, B)
    assert np.array_equal(crp, r), name + ": row pointers differ from cuSPARSE"
    # cuSPARSE does not promise sorted columns inside a row: compare row by row as sorted (column, value) lists
    order = np.lexsort((cci, np.repeat(np.arange(m), np.diff(crp))))
    assert np.array_equal(cci[order], c), name + ": structure differs from cuSPARSE"
    err = np.max(np.abs(cv[order] - v) / np.maximum(np.abs(v), 1e-300)) if len(v) else 0.0
    assert err <= 1e-12, (name, err)
