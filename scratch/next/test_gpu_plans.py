"""To be moved to tests/ once plans_integration.patch is applied and has passed on a B200 (round 2).
Runs the whole GPU parity suite a second time in a subprocess with TSG_PLANS=1 (the switch is read once per process), then
checks on two cases that the planned path really ran (launch count differs from the generic path) and that an R-MAT falls back."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))


def test_parity_suite_under_plans():
    env = dict(os.environ, TSG_PLANS="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x", "-m", "gpu"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


_PROBE = r"""
import sys, os, json
sys.path.insert(0, os.getcwd())
from spgemm_b200 import api, matrices as M
api.init(0)
out = {}
for name, gen in (("stencil", lambda: M.stencil27(24)), ("rmat", lambda: M.rmat(13, 16, seed=1))):
    m, n, rp, ci, v = gen()
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    c, st = api.spgemm(tA, tB)
    out[name] = [st["launches"], st["nnzC"], float(api.tile_rowsums(c)[0].sum())]
print(json.dumps(out))
"""


def test_planned_path_is_taken_on_structured_input_only():
    res = {}
    for plans in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", _PROBE], cwd=ROOT, env=dict(os.environ, TSG_PLANS=plans), capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        import json
        res[plans] = json.loads(r.stdout.strip().splitlines()[-1])
    for name in ("stencil", "rmat"):
        assert res["0"][name][1] == res["1"][name][1] and np.isclose(res["0"][name][2], res["1"][name][2], rtol=1e-12)
    assert res["1"]["stencil"][0] != res["0"]["stencil"][0], "plans were not used on the stencil"
