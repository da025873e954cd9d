#!/usr/bin/env python
"""bench.py -- TileSpGEMM hot path on B200: SpGEMM GFLOP/s (2 * nnzCub / time), FP64, 16x16 tiles.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of SpGEMM steps 1-3 (tile-level symbolic, bitmask symbolic, numeric) over the
workload with the tiled A and B already resident in HBM (`value`). `e2e` is the same metric measured
through the public API from HOST buffers: H2D of CSR(A) from pinned memory, csr2tile of A and B,
steps 1-3, tile2csr, D2H of CSR(C) into pinned memory, all inside the timed region (C leaves the
device slab by slab, the copy of one slab overlapping the computation of the next: tsg_spgemm_to_host).

N > 1 (spgemm_b200.multigpu): B is broadcast once over NCCL (as the CSR it is built from, or tiled with
--bcast tiled); the step-1 weights are computed in parallel and all-gathered; C tile-rows are partitioned
across the ranks (contiguous ranges of equal weight); each rank computes its C tile-rows with no further
communication. Total work is fixed => "scaling": "strong". Every line carries "parity": the run's own
C*ones and per-row counts, from the device, against the CPU.

--impl reference: the reference's own CPU path (oracle/_ref: unmodified spgemm_spa of
src/spgemm_serialref_spa_new.h, OpenMP on all host cores) on a bounded row sample of the same
workload, same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from spgemm_b200 import matrices as M  # noqa: E402

WORKLOADS = {
    # name: (generator, aat, description)  -- BASELINE.json configs
    "lap2d-256": (lambda: M.lap2d(256), False, "C=A^2, 2D 5-point Laplacian 256x256 (config 1)"),
    "stencil27-128": (lambda: M.stencil27(128), False, "C=A^2, 3D 27-point stencil 128^3 (config 2)"),
    "stencil27-64": (lambda: M.stencil27(64), False, "C=A^2, 3D 27-point stencil 64^3 (reduced, for quick checks)"),
    "rmat-s16-aat": (lambda: M.rmat(16, 16, seed=1), True, "C=AA^T, R-MAT scale 16 Graph500 skew (reduced config 3)"),
    "rmat-s18-aat": (lambda: M.rmat(18, 16, seed=1), True, "C=AA^T, R-MAT scale 18 Graph500 skew (reduced config 3)"),
    "rmat-s20-aat": (lambda: M.rmat(20, 16, seed=1), True, "C=AA^T, R-MAT scale 20, edge factor 16, Graph500 skew (0.57,0.19,0.19,0.05), slab-wise (config 3)"),
    "rmat-s22": (lambda: M.rmat(22, 16, a=0.30, b=0.25, c=0.25, d=0.20, seed=1), False,
                 "C=A^2, R-MAT scale 22, edge factor 16, mild skew (0.30,0.25,0.25,0.20), slab-wise (reduced config 5)"),
    "rmat-s24": (lambda: M.rmat(24, 16, a=0.30, b=0.25, c=0.25, d=0.20, seed=1), False,
                 "C=A^2, R-MAT scale 24, edge factor 16, mild skew (0.30,0.25,0.25,0.20), slab-wise (config 5)"),
    "blockfem-2M": (lambda: M.blockfem(333334), False, "C=A^2, block-FEM 2M rows, dense 6x6 blocks, band 1 (config 4)"),
    "mixed-fem-stencil": (lambda: mixed_fem_stencil(), False,
                          "C=A^2, block-diagonal [block-FEM 1M rows | 27-point stencil 100^3]: well-filled and sparse tiles in one matrix "
                          "(per-tile accumulator selection; not a BASELINE config)"),
}


def mixed_fem_stencil():
    """Block-diagonal of a block-FEM matrix and a 27-point stencil: half of the C tiles want the dense accumulator, half the sparse one."""
    import scipy.sparse as sp
    parts = []
    for m, n, rp, ci, v in (M.blockfem(166667), M.stencil27(100)):
        parts.append(sp.csr_matrix((np.ones(len(ci)), ci, rp), shape=(m, n)))
    S = sp.block_diag(parts, format="csr")
    S.sort_indices()
    return S.shape[0], S.shape[1], S.indptr.astype(np.int32), S.indices.astype(np.int32), M.set_values(S.nnz, "mod10")
# workloads whose C does not fit one GPU / int32 offsets whole: executed as slabs of at most this many tile pairs
SLAB_PAIRS = {"rmat-s16-aat": 1 << 26, "rmat-s18-aat": 1 << 28, "rmat-s20-aat": 1 << 28, "rmat-s22": 1 << 28, "rmat-s24": 1 << 28}
DEFAULT_WORKLOAD = "stencil27-128"   # BASELINE.json configs[1]: the configuration the metric is quoted on


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(sm)}


def step3_bytes(tA, tB, st):
    """Algorithmic HBM bytes of the numeric step: what it has to read and write once -- A and B tile
    payloads (Val 8 per nnz, Ptr 32 + mask 32 + tile_nnz 4 per tile), the pair lists (8 per pair), C's structure for
    the tiles that hold entries (Ptr 32 + mask 32 + tile_nnz 4 each: the empty listed tiles of a hypersparse product are
    never read by it) and C's payload written (Val 8 + Col 2 per nnz)."""
    if st.get("plan_recipes", 0) > 0:  # plan path (csrc/plans.cu): values of A and B, per pair its two tile indices (8),
        # per C tile its recipe id, nnz offset and pair offset (12), C's payload written; the plans themselves stay in L1/L2
        return tA.nnz * 8 + tB.nnz * 8 + st["pairs"] * 8 + st["numblkC"] * 12 + st["nnzC"] * 10
    return (tA.nnz * 8 + tA.numtile * 68 + tB.nnz * 8 + tB.numtile * 68 + st["tiles_nonempty"] * 68 + st["pairs"] * 8
            + st["nnzC"] * 10)


def numeric_kernels(st):
    """Names of the numeric (step 3) kernels this workload actually launched (tsg_stats)."""
    names = []
    if st.get("plan_recipes", 0) > 0:
        if st.get("row_templates", 0) > 0 and os.environ.get("TSG_ROWPLAN_NUMERIC", "1") != "0" and st["row_templates"] <= 256:
            return f"k_numeric_from_rowplans ({st['plan_recipes']} recipes, {st['row_templates']} tile-row templates)"
        return f"k_numeric_from_plans_rows ({st['plan_recipes']} recipes)"
    if st.get("rows_staged", 0) > 0:
        names.append(f"k_step3_rows ({st['rows_staged']} tile-rows, {st['rows_smem']} B smem)")
    if st.get("tiles_dense", 0) > 0:
        names.append(f"k_step3_dense ({st['tiles_dense']} tiles)")
    if st.get("rows_gather", 0) > 0:
        names.append(f"k_step3_gather ({st['rows_gather']} tile-rows)")
    return " + ".join(names) if names else "none"


def run_reference(args):
    """Reference CPU arm: unmodified spgemm_spa (two-pass protocol) from oracle/_ref on a row sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc, ref
    orc.set_num_threads()  # torchrun exports OMP_NUM_THREADS=1 to its workers: take every host core back
    gen, aat, desc = WORKLOADS[args.workload]
    m, n, rp, ci, v = gen()
    A = (rp, ci, v)
    B = A
    if aat:
        cp, ri, cv = orc.transpose(m, n, rp, ci, v)
        B = (cp.astype(np.int32), ri, cv)
    nB = m if aat else n
    R = min(m, args.ref_rows)
    sample = (rp[:R + 1], ci[:rp[R]], v[:rp[R]])
    products = orc.nnzcub(sample[1], B[0])
    total_products = orc.nnzcub(ci, B[0])
    kind = "reference" if ref.available() else "port"
    fn = (lambda: ref.spgemm_spa(sample, B, nB)) if kind == "reference" else (lambda: orc.spgemm_spa(sample, B, nB))
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    gf = 2.0 * products / dt / 1e9
    cores = orc.num_threads()
    sample_desc = (f"rows [0,{R}) of A ({products} products of the workload's {total_products}) x whole B; "
                   + ("reference spgemm_spa (src/spgemm_serialref_spa_new.h, structure-only, count+fill passes)"
                      if kind == "reference" else "oracle SPA with values"))
    line = {"impl": "reference", "metric": "spgemm_gflops", "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "tile": "16x16", "aat": int(aat), "m": m, "n": n,
                       "nnzA": int(len(ci)), "nnzCub": int(total_products), "sample_fraction": products / max(total_products, 1)},
            "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample_desc},
            "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(A, B, nB, total_products, budget_products=1.6e9):
    """Oracle SPA with values ("port") on all host cores, on a bounded sample of the same workload: whole A when it holds
    at most `budget_products` products, else a seeded random sample of A's rows (x whole B) holding about that many --
    rows drawn at random, not the first ones: R-MAT's heavy rows sit at the low indices."""
    from oracle import oracle as orc
    orc.set_num_threads()
    rp, ci, v = A
    m = len(rp) - 1
    if total_products <= budget_products:
        sample, what = A, f"all {m} rows of A"
    else:
        k = max(int(m * budget_products / total_products), 1)
        rows = np.sort(np.random.default_rng(12345).choice(m, size=k, replace=False))
        cnt = (rp[rows + 1] - rp[rows]).astype(np.int64)
        srp = np.concatenate([[0], np.cumsum(cnt)])
        idx = np.repeat(rp[rows].astype(np.int64) - srp[:-1], cnt) + np.arange(srp[-1])
        sample, what = (srp, ci[idx], v[idx]), f"{k} of {m} rows of A drawn at random (seed 12345)"
    products = orc.nnzcub(sample[1], B[0])
    t0 = time.perf_counter()
    orc.spgemm_spa(sample, B, nB)
    dt = time.perf_counter() - t0
    return {"value": 2.0 * products / dt / 1e9, "unit": "GFLOP/s", "cores": orc.num_threads(), "kind": "port",
            "sample": f"{what} x whole B, {products} products, {dt:.2f} s wall, OpenMP SPA with values (oracle/spa_ref.c)"}


def parity_check(A, B, nB, sums, counts, want_counts):
    """Size-independent check of a whole run against the CPU: C * ones = A * (B * ones) (exact in FP64 for the driver's
    integer-valued inputs: every partial sum is an integer below 2^53) and, when asked for, nnz of every row of C against
    the count pass of the serial SPA (oracle; the reference's get_nnzC_only protocol)."""
    import scipy.sparse as sp
    mA, mB = len(A[0]) - 1, len(B[0]) - 1
    SA = sp.csr_matrix((A[2], A[1], A[0]), shape=(mA, mB))
    SB = sp.csr_matrix((B[2], B[1], B[0]), shape=(mB, nB))
    exp = SA @ (SB @ np.ones(nB))
    out = {"rows": int(mA), "rowsums_equal": bool(np.array_equal(sums, exp)),
           "rowsums_max_rel_err": float(np.max(np.abs(sums - exp) / np.maximum(np.abs(exp), 1.0))) if mA else 0.0,
           "nnzC": int(counts.sum())}
    if want_counts:
        from oracle import oracle as orc
        orc.set_num_threads()
        t0 = time.perf_counter()
        ec = orc.spgemm_rowcounts(A, B, nB)
        out.update({"rowcounts_equal": bool(np.array_equal(counts, ec)), "oracle_nnzC": int(ec.sum()),
                    "oracle_count_pass_s": round(time.perf_counter() - t0, 2)})
    return out


def setup_nccl_logging(world):
    """NCCL's INFO log (topology, NVLS, 'comm ... nranks N') goes to a file per rank under gpurun_out/ so that stdout stays
    the one JSON line; whatever the caller already set in the environment is left alone."""
    if world <= 1:
        return None
    if "NCCL_DEBUG" in os.environ:
        return os.environ.get("NCCL_DEBUG_FILE")
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    os.environ["NCCL_DEBUG"] = "INFO"
    os.environ["NCCL_DEBUG_SUBSYS"] = "INIT,ENV"
    os.environ["NCCL_DEBUG_FILE"] = os.path.join(d, f"nccl_{world}gpu_%h_%p.log")  # NCCL expands %h (host) and %p (pid)
    return os.environ["NCCL_DEBUG_FILE"]


def nccl_evidence(path_pattern):
    """What the NCCL log of this run says about the communicator (rank 0's view): ranks, NVLS, channels."""
    import glob
    import re
    if not path_pattern:
        return None
    files = glob.glob(path_pattern.replace("%h", "*").replace("%p", "*"))
    nranks, nvls, lines = None, False, 0
    for f in files:
        try:
            for ln in open(f, errors="ignore"):
                lines += 1
                mo = re.search(r"nranks (\d+)", ln)
                if mo:
                    nranks = max(nranks or 0, int(mo.group(1)))
                if "NVLS" in ln:
                    nvls = True
        except OSError:
            pass
    return {"log_files": len(files), "log_lines": lines, "nranks": nranks, "nvls_mentioned": nvls}


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nccl_log = setup_nccl_logging(world)
    import torch
    from spgemm_b200 import api, multigpu as mg

    dist = None
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    api.init(local)

    gen, aat, desc = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup

    # ------------------------------------------------------------------ data: rank 0 generates, B is broadcast once
    A_host = gen() if rank == 0 else None
    sh = mg.distribute(A_host, aat, dist, dev, mode=args.bcast)
    m, n, nB, nnzCub = sh.m, sh.n, sh.nB, sh.nnzCub
    tA, tB, dA = sh.tA, sh.tB, sh.dA
    part = {"parts": [[int(a), int(b)] for a, b in zip(sh.cuts[:-1], sh.cuts[1:])], "imbalance": sh.imbalance}

    # ------------------------------------------------------------------ timed region: steps 1-3, inputs resident
    def barrier():
        api.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    slab_pairs = SLAB_PAIRS.get(args.workload)
    slab_w = api.tilerow_weights(tA, tB) if slab_pairs else None

    def spgemm_step(sink=None):
        """One pass of steps 1-3 over this rank's C tile-rows (slab by slab when C cannot be held whole)."""
        tot, _ = mg.spgemm(sh, slab_pairs, sink, weights=slab_w)
        return tot

    stats = []
    for _ in range(W):
        spgemm_step()
    barrier()
    launches0 = api.launch_count()
    with ClockSampler(local) as clk:
        api.timer_start()
        t0 = time.perf_counter()
        for _ in range(K):
            stats.append(spgemm_step())
        dev_ms = api.timer_stop()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
    gpu_launches = api.launch_count() - launches0
    clocks = clk.summary()

    # ------------------------------------------------------------------ parity of this very run (outside the timed region):
    # per-row sums and counts of C from the device, slab by slab, against the CPU on rank 0
    rows_local = dA.m
    sums, cnts = np.zeros(rows_local), np.zeros(rows_local, np.int64)
    if not args.no_parity:
        def check_sink(tC, st):
            r0 = tC.trow0 * 16
            s_, c_ = api.tile_rowsums(tC)
            sums[r0:r0 + len(s_)] = s_
            cnts[r0:r0 + len(c_)] = c_
        spgemm_step(check_sink)

    # ------------------------------------------------------------------ e2e: host CSR in -> host CSR out
    # (per rank: H2D of its CSR(A) rows, csr2tile(A); rank-local csr2tile(B) from its resident CSR is replaced,
    #  for N > 1, by the already broadcast tiled B -- the broadcast time is reported separately)
    rpA, ciA, vA = dA.download()
    pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (rpA, ciA, vA)]
    nnzC_local = stats[-1]["nnzC"]
    if slab_pairs:  # the pinned landing buffer is reused slab after slab: size it for the largest slab
        _, per = mg.spgemm(sh, slab_pairs, None, weights=slab_w)
        buf_nnz = max(p["nnzC"] for p in per)
    else:
        buf_nnz = nnzC_local
    out_pin = [torch.empty(dA.m + 1, dtype=torch.int32).pin_memory(), torch.empty(max(buf_nnz, 1), dtype=torch.int32).pin_memory(),
               torch.empty(max(buf_nnz, 1), dtype=torch.float64).pin_memory()]

    def land(tc, _st=None):
        cc = api.tile2csr_device(tc)
        cc.download_into(out_pin[0].data_ptr(), out_pin[1].data_ptr(), out_pin[2].data_ptr())
        cc.free()

    def e2e_step():
        a = api.DeviceCSR.upload_ptrs(dA.m, n, pin[0].data_ptr(), pin[1].data_ptr(), pin[2].data_ptr())
        ta = api.csr2tile(a, False)
        if world == 1:
            b = api.transpose(a) if aat else a
            tb = api.csr2tile(b, True)
        else:
            b, tb = a, tB
        if slab_pairs:
            api.spgemm_slabs(ta, tb, max_pairs=slab_pairs, sink=land)
        else:  # C leaves the device slab by slab, each copy overlapping the next slab's computation
            api.spgemm_to_host(ta, tb, out_pin[0].data_ptr(), out_pin[1].data_ptr(), out_pin[2].data_ptr(), buf_nnz,
                               nslabs=args.e2e_slabs)
        ta.free()
        if world == 1:
            tb.free()
            if aat:
                b.free()
        a.free()

    e2e_K = max(1, min(K, args.e2e_steps))
    for _ in range(max(1, min(W, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_K):
        e2e_step()
    api.sync()
    e2e_local_ms = (time.perf_counter() - t0) * 1e3 / e2e_K
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_K
    h2d = sum(int(t.numel() * t.element_size()) for t in pin)
    d2h = (dA.m + 1) * 4 + nnzC_local * 12

    # ------------------------------------------------------------------ reduce over ranks
    ms_step = max(dev_ms, 0.0) / K
    last = stats[-1]
    vec = [ms_step, e2e_ms, float(gpu_launches), float(h2d), float(d2h), float(last["nnzC"]), float(last["numblkC"]),
           float(last["pairs"]), float(np.mean([s["ms_step1"] for s in stats])), float(np.mean([s["ms_step2"] for s in stats])),
           float(np.mean([s["ms_step3"] for s in stats])), float(np.mean([s["ms_alloc"] for s in stats])),
           float(last["algorithmic_bytes"]), float(step3_bytes(tA, tB, last)), wall_ms / K, e2e_local_ms,
           float(last["rows_staged"]), float(last["rows_gather"]), float(last["tiles_dense"]), float(last["rows_smem"]),
           float(last["plan_recipes"]), float(last.get("row_templates", 0))]
    if dist is not None:
        t = torch.tensor(vec, dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        allv = np.stack([x.cpu().numpy() for x in allv])
        # the checksums of every rank's rows, on rank 0
        longest = int(max(b - a for a, b in part["parts"])) * 16
        ts = torch.zeros(max(longest, 1), dtype=torch.float64, device=dev)
        tc = torch.zeros(max(longest, 1), dtype=torch.int64, device=dev)
        ts[:rows_local] = torch.from_numpy(sums).to(dev)
        tc[:rows_local] = torch.from_numpy(cnts).to(dev)
        gs, gc = [torch.zeros_like(ts) for _ in range(world)], [torch.zeros_like(tc) for _ in range(world)]
        dist.all_gather(gs, ts)
        dist.all_gather(gc, tc)
        if rank == 0:
            lens = [min(b * 16, m) - min(a * 16, m) for a, b in part["parts"]]
            sums = np.concatenate([g.cpu().numpy()[:k] for g, k in zip(gs, lens)])
            cnts = np.concatenate([g.cpu().numpy()[:k] for g, k in zip(gc, lens)])
    else:
        allv = np.asarray([vec])
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    ms_step_max, e2e_ms_max = float(allv[:, 0].max()), float(allv[:, 1].max())
    value = 2.0 * nnzCub / (ms_step_max * 1e6)
    peak, peak_src = measured_peak_gbs()
    # dominant kernel = the numeric step (step 3); per rank: bytes / its duration; report the slowest rank's
    slow = int(np.argmax(allv[:, 10]))
    s3_ms, s3_bytes = float(allv[slow, 10]), float(allv[slow, 13])
    achieved = s3_bytes / (s3_ms * 1e-3) / 1e9 if s3_ms > 0 else 0.0
    alg_total = float(allv[:, 12].sum())  # every rank reads the whole B, so B's bytes count once per rank
    rp, ci, v = A_host[2], A_host[3], A_host[4]
    kern_stats = {"rows_staged": int(allv[:, 16].sum()), "rows_gather": int(allv[:, 17].sum()), "tiles_dense": int(allv[:, 18].sum()),
                  "rows_smem": int(allv[:, 19].max()), "plan_recipes": int(allv[:, 20].max()), "row_templates": int(allv[:, 21].max())}
    line = {
        "metric": "spgemm_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step_max, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "tile": "16x16", "aat": int(aat), "m": m, "n": n,
                   "nnzA": int(len(ci)), "nnzCub": int(nnzCub), "nnzC": int(allv[:, 5].sum()),
                   "C_tiles": int(allv[:, 6].sum()), "tile_pairs": int(allv[:, 7].sum()),
                   "plan_recipes": kern_stats["plan_recipes"], "row_templates": kern_stats["row_templates"],
                   "l2": "inputs larger than L2 (tiled A+B+C per step >> 126 MB)" if alg_total > 4 * 126e6 else
                         "working set fits L2: launch-latency-bound correctness config, not a roofline config",
                   "parallelism": f"tile-row partition x{world}", "partition": part,
                   "b_broadcast": {"mode": sh.mode, "bytes": sh.bcast_bytes, "ms": sh.bcast_ms,
                                   "gbs": (sh.bcast_bytes / sh.bcast_ms / 1e6) if sh.bcast_ms > 0 else None} if world > 1 else None,
                   "nccl": dict(nccl_evidence(nccl_log) or {}, NCCL_DEBUG=os.environ.get("NCCL_DEBUG"),
                                NCCL_DEBUG_FILE=os.environ.get("NCCL_DEBUG_FILE")) if world > 1 else None,
                   "steps_ms": {"step1": float(allv[:, 8].max()), "step2": float(allv[:, 9].max()), "step3": float(allv[:, 10].max()),
                                "alloc_and_sync": float(allv[:, 11].max()), "host_wall_per_step": float(allv[:, 14].max())},
                   "pipeline_roofline": {"algorithmic_bytes": alg_total, "achieved_gbs": alg_total / (ms_step_max * 1e-3) / 1e9,
                                         "frac_of_peak": alg_total / (ms_step_max * 1e-3) / 1e9 / (peak * world), "note": "SURVEY 8(d) bytes(A)+bytes(B)+bytes(C) over the whole step"}},
        "clocks": clocks,
        "e2e": {"value": 2.0 * nnzCub / (e2e_ms_max * 1e6), "unit": "GFLOP/s", "ms_per_step": e2e_ms_max, "steps": e2e_K,
                "h2d_bytes_per_step": int(allv[:, 3].sum()), "d2h_bytes_per_step": int(allv[:, 4].sum()),
                "per_rank_ms": [round(float(x), 3) for x in allv[:, 15]],
                "pcie_gbs_aggregate": (float(allv[:, 3].sum()) + float(allv[:, 4].sum())) / (e2e_ms_max * 1e-3) / 1e9},
        "gpu_launches": int(allv[:, 2].sum()),
        "roofline": {"bound": "hbm", "kernel": "numeric step: " + numeric_kernels(kern_stats), "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "ms_per_launch": s3_ms,
                     "algorithmic_bytes_per_launch": s3_bytes},
    }
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            line["roofline"]["traffic"] = json.load(open(prof)).get(args.workload)
        except Exception:
            pass
    if world == 1 or not args.no_parity or not args.no_cpu_baseline:
        from oracle import oracle as orc
        A = (rp, ci, v)
        B = A
        if aat:
            cp, ri, cv = orc.transpose(m, n, rp, ci, v)
            B = (cp.astype(np.int64), ri, cv)
        if not args.no_parity:
            line["parity"] = parity_check(A, B, nB, sums, cnts, args.parity_counts or nnzCub <= 4e9)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(A, B, nB, nnzCub)
    print(json.dumps(line), flush=True)
    sh.free()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_general_tiles(args):
    """--tile M N with (M, N) != (16, 16): the general-tile path (csrc/gentile.cu) on ONE GPU, same
    metric. Tiles of A are M x N, of B N x M, of C M x M (reference src/main.cu:84-91). `value`: steps 1-3 with the tiled A
    and B resident; `e2e`: pinned host CSR -> csr2tile x2 -> steps 1-3 -> tile2csr -> host CSR(C). Parity: the CSR(C) the
    end-to-end leg brought back, against the CPU (C*ones exact, nnz of every row against the oracle's count pass)."""
    import torch
    from spgemm_b200 import api
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    assert int(os.environ.get("WORLD_SIZE", "1")) == 1, "--tile other than 16 16 runs on one GPU"
    tm, tn = args.tile
    torch.cuda.set_device(0)
    api.init(0)
    gen, aat, desc = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup
    m, n, rp, ci, v = gen()
    dA = api.DeviceCSR.upload(m, n, rp, ci, v)
    dB = api.transpose(dA) if aat else dA
    nB = m if aat else n
    nnzCub = api.nnzcub(dA, dB)
    tA, tB = api.gtile_csr2tile(dA, False, tm, tn), api.gtile_csr2tile(dB, True, tn, tm)

    def step():
        c, st = api.gtile_spgemm(tA, tB)
        c.free()
        return st

    for _ in range(W):
        step()
    api.sync()
    stats = []
    launches0 = api.launch_count()
    with ClockSampler(0) as clk:
        api.timer_start()
        for _ in range(K):
            stats.append(step())
        dev_ms = api.timer_stop()
    gpu_launches = api.launch_count() - launches0
    clocks = clk.summary()
    last = stats[-1]

    # end to end from pinned host buffers
    pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (rp, ci, v)]
    nnzC = int(last["nnzC"])
    out_pin = [torch.empty(m + 1, dtype=torch.int32).pin_memory(), torch.empty(max(nnzC, 1), dtype=torch.int32).pin_memory(),
               torch.empty(max(nnzC, 1), dtype=torch.float64).pin_memory()]

    def e2e_step():
        a = api.DeviceCSR.upload_ptrs(m, n, pin[0].data_ptr(), pin[1].data_ptr(), pin[2].data_ptr())
        b = api.transpose(a) if aat else a
        ta, tb = api.gtile_csr2tile(a, False, tm, tn), api.gtile_csr2tile(b, True, tn, tm)
        tc, _ = api.gtile_spgemm(ta, tb)
        cc = api.gtile_tile2csr(tc)
        cc.download_into(out_pin[0].data_ptr(), out_pin[1].data_ptr(), out_pin[2].data_ptr())
        for o in (cc, tc, ta, tb) + ((b,) if aat else ()) + (a,):
            o.free()

    e2e_K = max(1, min(K, args.e2e_steps))
    e2e_step()
    api.sync()
    t0 = time.perf_counter()
    for _ in range(e2e_K):
        e2e_step()
    api.sync()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_K

    ms_step = dev_ms / K
    s3_ms = float(np.mean([s["ms_step3"] for s in stats]))
    W16 = tm // 16
    # algorithmic bytes of the numeric kernel (k_g_numeric): A's payload and row offsets, B's values, row offsets and row masks,
    # the pair lists, C's row offsets / masks / tile offsets read and C's payload written
    s3_bytes = (tA.nnz * 10 + tA.numtile * (2 * tm + 4) + tB.nnz * 8 + tB.numtile * (2 * tn + 2 * tn * W16 + 4) + last["pairs"] * 8
                + last["numblkC"] * (2 * tm + 2 * tm * W16 + 8) + nnzC * 10)
    peak, peak_src = measured_peak_gbs()
    achieved = s3_bytes / (s3_ms * 1e-3) / 1e9 if s3_ms > 0 else 0.0
    line = {
        "metric": "spgemm_gflops", "value": 2.0 * nnzCub / (ms_step * 1e6), "unit": "GFLOP/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "tile": f"{tm}x{tn}", "tiles_of": {"A": [tm, tn], "B": [tn, tm], "C": [tm, tm]},
                   "path": "general tiles (csrc/gentile.cu)", "aat": int(aat), "m": m, "n": n, "nnzA": int(len(ci)), "nnzCub": int(nnzCub),
                   "nnzC": nnzC, "A_tiles": int(tA.numtile), "B_tiles": int(tB.numtile), "C_tiles": int(last["numblkC"]),
                   "tile_pairs": int(last["pairs"]),
                   "l2": "inputs larger than L2" if last["algorithmic_bytes"] > 4 * 126e6 else "working set fits L2",
                   "steps_ms": {"step1": float(np.mean([s["ms_step1"] for s in stats])), "step2": float(np.mean([s["ms_step2"] for s in stats])),
                                "step3": s3_ms},
                   "pipeline_roofline": {"algorithmic_bytes": int(last["algorithmic_bytes"]),
                                         "achieved_gbs": last["algorithmic_bytes"] / (ms_step * 1e-3) / 1e9,
                                         "frac_of_peak": last["algorithmic_bytes"] / (ms_step * 1e-3) / 1e9 / peak}},
        "clocks": clocks,
        "e2e": {"value": 2.0 * nnzCub / (e2e_ms * 1e6), "unit": "GFLOP/s", "ms_per_step": e2e_ms, "steps": e2e_K,
                "h2d_bytes_per_step": sum(int(t.numel() * t.element_size()) for t in pin), "d2h_bytes_per_step": (m + 1) * 4 + nnzC * 12},
        "gpu_launches": int(gpu_launches),
        "roofline": {"bound": "hbm", "kernel": "numeric step: " + ("k_g_numeric_dense32 (warp per 32x32 C tile, register accumulators)" if last.get("tiles_dense", 0) > 0
                                                 else "k_g_numeric (thread per C nonzero)"), "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "ms_per_launch": s3_ms,
                     "algorithmic_bytes_per_launch": int(s3_bytes)},
    }
    if not args.no_parity or not args.no_cpu_baseline:
        from oracle import oracle as orc
        A = (rp, ci, v)
        B = A
        if aat:
            cp, ri, cv = orc.transpose(m, n, rp, ci, v)
            B = (cp.astype(np.int64), ri, cv)
        if not args.no_parity:
            crp = out_pin[0].numpy().astype(np.int64)
            cvv = out_pin[2].numpy()[:nnzC]
            cnts = np.diff(crp)
            csum = np.concatenate([[0.0], np.cumsum(cvv)])
            # row sums in one pass (exact: every partial sum of the driver's integer-valued inputs is an integer below 2^53)
            sums = csum[crp[1:]] - csum[crp[:-1]]
            line["parity"] = parity_check(A, B, nB, sums, cnts, args.parity_counts or nnzCub <= 4e9)
            cols = out_pin[1].numpy()[:nnzC]
            drops = np.flatnonzero(np.diff(cols) <= 0) + 1   # a column index may only fall where a new row starts
            line["parity"]["columns_sorted_in_rows"] = bool(np.all(np.isin(drops, crp)))
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(A, B, nB, nnzCub)
    print(json.dumps(line), flush=True)
    for o in (tA, tB) + ((dB,) if aat else ()) + (dA,):
        o.free()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=3, help="timed end-to-end steps (each moves GBs over PCIe)")
    ap.add_argument("--e2e-slabs", type=int, default=0, help="slabs of the overlapped end-to-end call (0 = library default)")
    ap.add_argument("--ref-rows", type=int, default=1 << 17, help="--impl reference: rows of A in the bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the C*ones / row-count check of the run against the CPU")
    ap.add_argument("--parity-counts", action="store_true",
                    help="also compare nnz per row of C with the oracle's count pass on workloads above 4e9 products (minutes of CPU)")
    ap.add_argument("--bcast", default="csr", choices=["csr", "tiled"], help="N > 1: broadcast B as its CSR (default) or as the tiled matrix")
    ap.add_argument("--tile", type=int, nargs=2, default=[16, 16], metavar=("M", "N"),
                    help="tile_size_m tile_size_n (multiples of 16 up to 128). 16 16: the tuned kernels (default, the BASELINE metric); "
                         "anything else: the general-tile path on one GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3  # timing rule: at least 3 warm-up steps
        if tuple(args.tile) != (16, 16):
            run_general_tiles(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
