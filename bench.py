#!/usr/bin/env python
"""bench.py -- TileSpGEMM hot path on B200: SpGEMM GFLOP/s (2 * nnzCub / time), FP64, 16x16 tiles.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of SpGEMM steps 1-3 (tile-level symbolic, bitmask symbolic, numeric) over the
workload with the tiled A and B already resident in HBM (`value`). `e2e` is the same metric measured
through the public API from HOST buffers: H2D of CSR(A) from pinned memory, csr2tile of A and B,
steps 1-3, tile2csr, D2H of CSR(C) into pinned memory, all inside the timed region (C leaves the
device slab by slab, the copy of one slab overlapping the computation of the next: tsg_spgemm_to_host).

N > 1: C tile-rows are partitioned across the ranks (contiguous ranges balanced by the step-1
weight); B is tiled once on rank 0 and broadcast as one buffer over NCCL; each rank computes its C
tile-rows with no further communication. Total work is fixed => "scaling": "strong".

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
}
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
    """Algorithmic HBM bytes of the numeric kernel: A and B tile payloads it reads
    (Val 8 + Col 2 per nnz, Ptr 32 + tile_nnz 4 per tile), C structure it reads (Ptr 32 + mask 32 +
    tile_nnz 4 per C tile, pair list 8 per pair) and C payload it writes (Val 8 + Col 2 per nnz)."""
    return (tA.nnz * 10 + tA.numtile * 36 + tB.nnz * 10 + tB.numtile * 36 + st["numblkC"] * 68 + st["pairs"] * 8
            + st["nnzC"] * 10)


def numeric_kernels(st):
    """Names of the numeric (step 3) kernels this workload actually launched (tsg_stats)."""
    names = []
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
    """Oracle SPA with values ("port") on all host cores, on a bounded sample of the same workload."""
    from oracle import oracle as orc
    rp = A[0]
    m = len(rp) - 1
    R = m
    if total_products > budget_products:  # first rows holding ~budget products
        R = max(int(m * budget_products / total_products), 1)
    sample = (rp[:R + 1], A[1][:rp[R]], A[2][:rp[R]])
    products = orc.nnzcub(sample[1], B[0])
    t0 = time.perf_counter()
    orc.spgemm_spa(sample, B, nB)
    dt = time.perf_counter() - t0
    return {"value": 2.0 * products / dt / 1e9, "unit": "GFLOP/s", "cores": orc.num_threads(), "kind": "port",
            "sample": f"rows [0,{R}) of A x whole B, {products} products, {dt:.2f} s wall, OpenMP SPA with values (oracle/spa_ref.c)"}


def run_ours(args):
    import torch
    from spgemm_b200 import api, multigpu as mg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there when NCCL_DEBUG is VERSION/INFO
        os.environ["NCCL_DEBUG"] = os.environ.get("TSG_NCCL_DEBUG", "WARN")
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    api.init(local)

    gen, aat, desc = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup

    # ------------------------------------------------------------------ data (rank 0 generates)
    bcast_ms = 0.0
    if rank == 0:
        m, n, rp, ci, v = gen()
        dA_full = api.DeviceCSR.upload(m, n, rp, ci, v)
        dB = api.transpose(dA_full) if aat else dA_full
        nnzCub = api.nnzcub(dA_full, dB)
        tB = api.csr2tile(dB, True)
        nB = dB.n
    if world == 1:
        dA, tA = dA_full, api.csr2tile(dA_full, False)
        part = {"parts": [[0, tA.tilem]], "imbalance": 1.0}
    else:
        # sizes of B, partition of A's tile-rows, per-rank CSR sizes
        if rank == 0:
            tA_full = api.csr2tile(dA_full, False)
            w = api.tilerow_weights(tA_full, tB)
            cuts = mg.partition_tilerows(w, world)
            tA_full.free()
            hdr = [tB.m, tB.n, tB.numtile, tB.nnz, m, n, int(nnzCub)] + [int(c) for c in cuts]
            part = {"parts": [[int(a), int(b)] for a, b in zip(cuts[:-1], cuts[1:])], "imbalance": mg.imbalance(w, cuts)}
        else:
            hdr = [0] * (7 + world + 1)
            part = None
        h = torch.tensor(hdr, dtype=torch.int64, device=dev)
        dist.broadcast(h, 0)
        hdr = [int(x) for x in h.cpu()]
        bm, bn, bnt, bnnz, m, n, nnzCub = hdr[:7]
        cuts = hdr[7:]
        nB = bn
        if rank != 0:
            tB = api.tile_alloc(bm, bn, bnt, bnnz, True)
        # B: one buffer, one NCCL broadcast over NVLink
        slab = mg.tile_slab_tensor(tB, dev)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(slab, 0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        bcast_bytes = slab.numel()
        # A: CSR rows of each rank's tile-row range (rank 0 slices on the host and sends)
        r0, r1 = cuts[rank] * 16, min(cuts[rank + 1] * 16, m)
        if rank == 0:
            for dst in range(1, world):
                a0, a1 = cuts[dst] * 16, min(cuts[dst + 1] * 16, m)
                srp, sci, sv = mg.csr_row_slice(rp, ci, v, a0, a1)
                dist.send(torch.tensor([len(sci)], dtype=torch.int64, device=dev), dst)
                for arr in (srp, sci, sv):
                    dist.send(torch.from_numpy(np.ascontiguousarray(arr)).to(dev), dst)
            srp, sci, sv = mg.csr_row_slice(rp, ci, v, r0, r1)
            dA = api.DeviceCSR.upload(r1 - r0, n, srp, sci, sv)
        else:
            cnt = torch.zeros(1, dtype=torch.int64, device=dev)
            dist.recv(cnt, 0)
            nz = int(cnt.item())
            t_rp = torch.empty(r1 - r0 + 1, dtype=torch.int32, device=dev)
            t_ci = torch.empty(max(nz, 1), dtype=torch.int32, device=dev)
            t_v = torch.empty(max(nz, 1), dtype=torch.float64, device=dev)
            for t in (t_rp, t_ci[:nz], t_v[:nz]):
                dist.recv(t, 0)
            torch.cuda.synchronize()
            recv_bufs = (t_rp, t_ci, t_v)  # keep the received tensors alive: dA borrows their memory
            dA = api.DeviceCSR.wrap(r1 - r0, n, nz, t_rp.data_ptr(), t_ci.data_ptr(), t_v.data_ptr())
        tA = api.csr2tile(dA, False)

    # ------------------------------------------------------------------ timed region: steps 1-3, inputs resident
    def barrier():
        api.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    slab_pairs = SLAB_PAIRS.get(args.workload)
    slab_w = api.tilerow_weights(tA, tB) if slab_pairs else None

    def spgemm_step():
        """One pass of steps 1-3 over this rank's C tile-rows (slab by slab when C cannot be held whole)."""
        if slab_pairs:
            tot, _ = api.spgemm_slabs(tA, tB, max_pairs=slab_pairs, weights=slab_w)
            return None, tot
        tC_, st_ = api.spgemm(tA, tB)
        tC_.free()
        return None, st_

    stats = []
    for _ in range(W):
        spgemm_step()
    barrier()
    launches0 = api.launch_count()
    with ClockSampler(local) as clk:
        api.timer_start()
        t0 = time.perf_counter()
        for _ in range(K):
            _, st = spgemm_step()
            stats.append(st)
        dev_ms = api.timer_stop()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
    gpu_launches = api.launch_count() - launches0
    clocks = clk.summary()

    # ------------------------------------------------------------------ e2e: host CSR in -> host CSR out
    # (per rank: H2D of its CSR(A) rows, csr2tile(A); rank-local csr2tile(B) from its resident CSR is replaced,
    #  for N > 1, by the already broadcast tiled B -- the broadcast time is reported separately)
    rpA, ciA, vA = dA.download() if world > 1 else (rp, ci, v)
    pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (rpA, ciA, vA)]
    nnzC_local = stats[-1]["nnzC"]
    if slab_pairs:  # the pinned landing buffer is reused slab after slab: size it for the largest slab
        _, per = api.spgemm_slabs(tA, tB, max_pairs=slab_pairs, weights=slab_w)
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
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_K
    h2d = sum(int(t.numel() * t.element_size()) for t in pin)
    d2h = (dA.m + 1) * 4 + nnzC_local * 12

    # ------------------------------------------------------------------ reduce over ranks
    ms_step = max(dev_ms, 0.0) / K
    vec = [ms_step, e2e_ms, float(gpu_launches), float(h2d), float(d2h), float(stats[-1]["nnzC"]), float(stats[-1]["numblkC"]),
           float(stats[-1]["pairs"]), float(np.mean([s["ms_step1"] for s in stats])), float(np.mean([s["ms_step2"] for s in stats])),
           float(np.mean([s["ms_step3"] for s in stats])), float(np.mean([s["ms_alloc"] for s in stats])),
           float(stats[-1]["algorithmic_bytes"]), float(step3_bytes(tA, tB, stats[-1])), wall_ms / K]
    if dist is not None:
        t = torch.tensor(vec, dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        allv = np.stack([x.cpu().numpy() for x in allv])
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
    # dominant kernel = the numeric kernel (step 3); per rank: bytes / its duration; report the slowest rank's kernel
    slow = int(np.argmax(allv[:, 10]))
    s3_ms, s3_bytes = float(allv[slow, 10]), float(allv[slow, 13])
    achieved = s3_bytes / (s3_ms * 1e-3) / 1e9 if s3_ms > 0 else 0.0
    alg_total = float(allv[:, 12].sum())  # every rank reads the whole B, so B's bytes count once per rank
    line = {
        "metric": "spgemm_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step_max, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "tile": "16x16", "aat": int(aat), "m": m, "n": n,
                   "nnzA": int(len(ci)) if world == 1 else None, "nnzCub": int(nnzCub), "nnzC": int(allv[:, 5].sum()),
                   "C_tiles": int(allv[:, 6].sum()), "tile_pairs": int(allv[:, 7].sum()),
                   "l2": "inputs larger than L2 (tiled A+B+C per step >> 126 MB)" if alg_total > 4 * 126e6 else
                         "working set fits L2: launch-latency-bound correctness config, not a roofline config",
                   "parallelism": f"tile-row partition x{world}", "partition": part,
                   "b_broadcast_ms": bcast_ms if world > 1 else None,
                   "b_broadcast_gbs": (bcast_bytes / bcast_ms / 1e6) if world > 1 and bcast_ms > 0 else None,
                   "steps_ms": {"step1": float(allv[:, 8].max()), "step2": float(allv[:, 9].max()), "step3": float(allv[:, 10].max()),
                                "alloc_and_sync": float(allv[:, 11].max()), "host_wall_per_step": float(allv[:, 14].max())},
                   "pipeline_roofline": {"algorithmic_bytes": alg_total, "achieved_gbs": alg_total / (ms_step_max * 1e-3) / 1e9,
                                         "frac_of_peak": alg_total / (ms_step_max * 1e-3) / 1e9 / (peak * world), "note": "SURVEY 8(d) bytes(A)+bytes(B)+bytes(C) over the whole step"}},
        "clocks": clocks,
        "e2e": {"value": 2.0 * nnzCub / (e2e_ms_max * 1e6), "unit": "GFLOP/s", "ms_per_step": e2e_ms_max, "steps": e2e_K,
                "h2d_bytes_per_step": int(allv[:, 3].sum()), "d2h_bytes_per_step": int(allv[:, 4].sum())},
        "gpu_launches": int(allv[:, 2].sum()),
        "roofline": {"bound": "hbm", "kernel": numeric_kernels(stats[-1]), "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "ms_per_launch": s3_ms,
                     "algorithmic_bytes_per_launch": s3_bytes},
    }
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            line["roofline"]["traffic"] = json.load(open(prof)).get(args.workload)
        except Exception:
            pass
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        A = (rp, ci, v)
        B = A
        if aat:
            cp, ri, cv = orc.transpose(m, n, rp, ci, v)
            B = (cp.astype(np.int32), ri, cv)
        line["cpu_baseline"] = cpu_baseline(A, B, nB, nnzCub)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3  # timing rule: at least 3 warm-up steps
        run_ours(args)


if __name__ == "__main__":
    main()
