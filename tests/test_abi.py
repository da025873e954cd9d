"""The C-ABI shared library loads and exports every symbol include/tilespgemm.h declares; the ctypes
mirrors match the C layouts; without a GPU the product fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT
from spgemm_b200 import lib as L


def header_functions():
    txt = open(os.path.join(ROOT, "include", "tilespgemm.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:unsigned\s+)?(?:void|int|char|long long)\s*\*?\s*(\w+)\s*\(", txt, flags=re.M)
    return sorted(set(names))


def test_header_matches_exports():
    hdr = header_functions()
    assert sorted(L.EXPORTS) == hdr, (set(hdr) ^ set(L.EXPORTS))
    lib = L.load()
    for name in hdr:
        assert hasattr(lib, name), f"{name} declared in include/tilespgemm.h but not exported"


def test_struct_layouts(tmp_path):
    """sizeof/offsetof from a C compile of the header against the ctypes mirrors."""
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "tilespgemm.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(SMatrix), sizeof(tsg_dcsr), sizeof(tsg_dtile),'
        ' sizeof(tsg_stats), offsetof(SMatrix, tile_csr_Ptr), offsetof(tsg_dtile, rm2csc), offsetof(tsg_dtile, slab_bytes),'
        ' offsetof(tsg_stats, launches), sizeof(tsg_gtile), offsetof(tsg_gtile, nnz), offsetof(tsg_gtile, slab_bytes));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    exp = [C.sizeof(L.SMatrix), C.sizeof(L.DCsr), C.sizeof(L.DTile), C.sizeof(L.Stats), L.SMatrix.tile_csr_Ptr.offset,
           L.DTile.rm2csc.offset, L.DTile.slab_bytes.offset, L.Stats.launches.offset, C.sizeof(L.GTile), L.GTile.nnz.offset,
           L.GTile.slab_bytes.offset]
    assert got == exp


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_fails_loudly_without_gpu():
    lib = L.load()
    lib.tilespgemm_clear_error()
    assert lib.tsg_init(0) == 1  # TSG_ERR_CUDA
    assert b"no CPU fallback" in lib.tilespgemm_last_error_string() or b"CUDA" in lib.tilespgemm_last_error_string()
    with pytest.raises(L.TsgError):
        L.check()
    assert lib.tilespgemm_last_error() == 0  # check() clears the latch
