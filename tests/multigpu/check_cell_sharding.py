"""Multi-GPU check of the cell-sharded solve (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port 29511 tests/multigpu/check_cell_sharding.py

Every rank solves the problem alone on its GPU, then all ranks solve it together
(a) over peer memory (remote stores from the cell solver + flag barrier) and
(b) with the NCCL all-gather path; all three must agree bit for bit (Jacobi) with
the same sweep count.  Prints per-sweep device times of (a) and (b) at C4 size.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rabacus_b200 import configs  # noqa: E402
from rabacus_b200.plan import FIELDS, REC_FIELDS, Plan  # noqa: E402


def make_plan(cfg, local_rank, **kw):
    Nl, Nnu = cfg["nH"].size, cfg["E_eV"].size
    plan = Plan(configs.GEOM_IDS[cfg["geometry"]], 1, Nl, Nnu, cfg["Nmu"],
                find_Teq=cfg["find_Teq"], fixed_fcA=cfg["fixed_fcA"], tol=1e-4,
                device=local_rank, i_rec_meth=cfg.get("i_rec_meth", 1),
                em_fac=cfg.get("em_fac", (1.0, 1.0, 1.0)), **kw)
    plan.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"],
                     cfg["shape"], cfg["z"], 0.0)
    plan.upload()
    return plan


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ok = True
    cases = [("sphere 256x128x16", configs.c4_sphere_bgnd(256, 128, 16, False, nH0=3.0)),
             ("sphere 256x128x16 Teq", configs.c4_sphere_bgnd(256, 128, 16, True, nH0=3.0)),
             ("slab2bgnd 512x128", configs.c3_slab_bgnd(512, 128, True)),
             ("slabpln 256x128", configs.c2_slab_plane(256, 128, False))]
    # recombination-photon transfer under cell sharding: every rank runs the closure for
    # its own solve layers only (rb_rec.cuh RecCells)
    for meth, tag in ((2, "outward"), (3, "radial"), (4, "isotropic")):
        cfg = configs.c4_sphere_bgnd(256, 128, 16, meth == 3, nH0=3.0)
        cfg["i_rec_meth"] = meth
        if meth == 4:
            cfg["em_fac"] = (1.0, 0.7, 1.3)
        cases.append(("sphere 256 rec %s" % tag, cfg))
    cfg = configs.c1_iliev_stromgren(256)
    cfg["i_rec_meth"] = 4
    cfg["Nmu"] = 8
    cases.append(("stromgren 256 rec iso", cfg))
    for nm, cfg in (("slab2bgnd 512 rec ray", configs.c3_slab_bgnd(512, 128, True)),
                    # odd Nl: the middle layer is never solved; Nl/2 divisible by the world size
                    ("slabpln2 odd rec ray", configs.c2_slab_plane(2 * world * 31 + 1, 128, False,
                                                                   True)),
                    ("slabpln 256 rec ray", configs.c2_slab_plane(256, 128, False))):
        cfg["i_rec_meth"] = 2
        cases.append((nm, cfg))
    for name, cfg in cases:
        p1 = make_plan(cfg, local_rank)
        i1 = p1.solve()
        ref = p1.get_problem(0)
        p1.close()
        res = {}
        for mode in ("peer", "nccl"):
            p = make_plan(cfg, local_rank)
            if mode == "peer":
                p.connect_peers(rank, world)
            n = p.solve_cell_sharded(rank, world)
            out = p.get_problem(0)
            p.close()
            same = n == i1["iters"] and all(np.array_equal(out[k], ref[k])
                                            for k in FIELDS + REC_FIELDS)
            res[mode] = (n, same)
            ok = ok and same
        if rank == 0:
            print("%-24s single %d sweeps | peer %s | nccl %s" % (name, i1["iters"], res["peer"],
                                                                  res["nccl"]), flush=True)
    # timing at C4 size: peer-memory path vs NCCL path, device ms per sweep
    cfg = configs.c4_sphere_bgnd(4096, 512, 256, False)
    for mode in (() if os.environ.get("RB_CHECK_NO_TIMING") else ("peer", "nccl")):
        p = make_plan(cfg, local_rank, max_sweeps=-(10 ** 9))
        torch.cuda.synchronize(); dist.barrier()
        if mode == "peer":
            p.connect_peers(rank, world)
            p.thin()
            p.bench_steps(3, 0)
            torch.cuda.synchronize(); dist.barrier()
            ms = p.bench_steps(20, 0) / 20
        else:
            send, recv = p.shard_buffers(world)
            ext = torch.cuda.ExternalStream(p.stream_ptr)
            with torch.cuda.stream(ext):
                p.thin()
                for _ in range(3):
                    p.sharded_sweep(rank, world, send, recv)
                torch.cuda.synchronize(); dist.barrier()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                ms = 0.0
                for _ in range(20):
                    e0.record(ext); p.sharded_sweep(rank, world, send, recv); e1.record(ext)
                    e1.synchronize(); ms += e0.elapsed_time(e1)
                ms /= 20
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("C4 sweep over %d GPUs, %s path: %.3f ms/sweep (max over ranks)"
                  % (world, mode, float(t.item())), flush=True)
        p.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        raise SystemExit("cell-sharding check FAILED")
    if rank == 0:
        print("cell-sharding check OK", flush=True)


if __name__ == "__main__":
    main()
