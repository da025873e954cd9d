"""torchrun worker of tests/test_multigpu_gpu.py: the multi-rank path of spgemm_b200.multigpu on real GPUs
(broadcast of B over NCCL, per-rank weights, device-side row slices, slab execution, gather of C) against the oracle.
Run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_worker.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    from oracle import oracle as orc
    from spgemm_b200 import api, matrices as M, multigpu as mg
    api.init(local)
    cases = [("stencil27_24 A^2, B as CSR", lambda: M.stencil27(24), False, None, "csr"),
             ("stencil27_20x9x7 A^2, B tiled", lambda: M.stencil27(20, 9, 7), False, None, "tiled"),
             ("rmat_s12 AA^T, slabs, B as CSR", lambda: M.rmat(12, 16, seed=6), True, 150000, "csr"),
             ("rmat_s11 A^2 hypersparse, B tiled", lambda: M.rmat(11, 3, a=0.30, b=0.25, c=0.25, d=0.20, seed=2), False, None, "tiled")]
    for name, gen, aat, slab_pairs, mode in cases:
        A_host = gen() if rank == 0 else None
        sh = mg.distribute(A_host, aat, dist, dev, mode=mode)
        assert sh.world == world and sh.cuts[0] == 0 and sh.cuts[-1] == (sh.m + 15) // 16
        pieces = []

        def sink(tC, st):
            csr = api.tile2csr_device(tC)
            pieces.append(csr.download())
            csr.free()

        tot, per = mg.spgemm(sh, slab_pairs, sink)
        if slab_pairs:
            assert len(per) >= 2 or world > 2, "the case must run more than one slab per rank"
        # this rank's CSR: its slabs one after the other
        rp = np.concatenate([[0]] + [p[0][1:] + off for p, off in zip(pieces, np.cumsum([0] + [len(p[1]) for p in pieces[:-1]]))])
        ci = np.concatenate([p[1] for p in pieces]) if pieces else np.zeros(0, np.int32)
        vv = np.concatenate([p[2] for p in pieces]) if pieces else np.zeros(0)
        rows = min(sh.trow1 * 16, sh.m) - min(sh.trow0 * 16, sh.m)
        assert len(rp) == rows + 1 and rp[-1] == tot["nnzC"] == len(ci), (name, len(rp), rows, rp[-1], tot["nnzC"])
        local_csr = api.DeviceCSR.upload(rows, sh.nB, rp.astype(np.int32), ci, vv)
        off = mg.concat(sh, tot, dist, dev)
        full = mg.gather_csr(sh, local_csr, dist, dev)
        local_csr.free()
        if rank == 0:
            m, n, arp, aci, av = A_host
            A = (arp, aci, av)
            B = A
            if aat:
                cp, ri, cv = orc.transpose(m, n, arp, aci, av)
                B = (cp, ri, cv)
            er, ec, ev = orc.spgemm_spa(A, B, sh.nB)
            assert off["nnzC"] == er[-1] and off["tile_offset"] == 0 and off["nnz_offset"] == 0, name
            assert np.array_equal(full[0], er), name + ": rowptr"
            assert np.array_equal(full[1], ec), name + ": colidx"
            assert np.array_equal(full[2], ev), name + ": values (integer-valued inputs must be bit-exact)"
            assert sh.nnzCub == orc.nnzcub(aci, B[0]), name
            print(f"ok: {name}: {world} ranks, cuts {sh.cuts.tolist()}, imbalance {sh.imbalance:.3f}, nnzC {er[-1]}, "
                  f"B broadcast {sh.bcast_bytes} B in {sh.bcast_ms:.3f} ms ({mode})", flush=True)
        sh.free()
        dist.barrier()
    # the one-call form (multigpu.run = SURVEY 8b's tilespgemm_multi): A A^T in slabs, whole CSR(C) gathered on rank 0
    A_host = M.stencil27(14, 9, 11) if rank == 0 else None
    res = mg.run(A_host, True, dist, dev, slab_pairs=60000, gather=True)
    assert res["offsets"]["row_offset"] == res["row0"] and len(res["local_csr"][0]) == res["row1"] - res["row0"] + 1
    if rank == 0:
        m, n, arp, aci, av = A_host
        B = orc.transpose(m, n, arp, aci, av)
        er, ec, ev = orc.spgemm_spa((arp, aci, av), B, m)
        assert np.array_equal(res["csr"][0], er) and np.array_equal(res["csr"][1], ec) and np.array_equal(res["csr"][2], ev), "run()"
        assert res["offsets"]["nnzC"] == er[-1] and res["nnzCub"] == orc.nnzcub(aci, B[0])
        print(f"ok: multigpu.run AA^T: {world} ranks, nnzC {er[-1]}", flush=True)
    else:
        assert res["csr"] is None
    dist.barrier()
    if rank == 0:
        print("MGPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
