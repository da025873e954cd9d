"""bench.py's reference arm (runs on the host cores, no GPU): one JSON line with the keys the driver reads, timed on the
reference's own unmodified CPU code (oracle/_ref, built from /root/reference/src/spgemm_serialref_spa_new.h) or, when that
build is absent, on the oracle port; and the product arm must refuse to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, env=e, timeout=600)


@pytest.mark.parametrize("omp", [None, "1"])
def test_reference_arm_prints_one_contract_line(omp):
    # torchrun exports OMP_NUM_THREADS=1; the arm must use all host cores regardless
    p = _run("--impl", "reference", "--workload", "lap2d-256", "--steps", "2", "--warmup", "1", env={"OMP_NUM_THREADS": omp} if omp else None)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "spgemm_gflops" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "lap2d-256" and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"] and cb["sample"]
    assert cb["cores"] >= 1
    if (os.cpu_count() or 1) > 1:
        assert cb["cores"] > 1, "the reference arm must not run single-threaded because OMP_NUM_THREADS=1 was exported"
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a box without a GPU")
def test_product_arm_fails_loudly_without_a_gpu():
    p = _run("--workload", "lap2d-256", "--steps", "1", "--warmup", "3", "--no-cpu-baseline")
    assert p.returncode != 0, "bench.py must not produce a number without the CUDA path"
    assert not [ln for ln in p.stdout.splitlines() if ln.strip().startswith("{")], p.stdout
