"""Soak of the cell-sharded solve over peer memory (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 tests/multigpu/soak_peer.py [--solves 24]

The same connected plan is solved `--solves` times back to back (~1000 sweeps at the default,
every one crossing the flag barrier and the remote-store exchange), alternating fixed T and
find_Teq state through re-uploads; after every solve the result must equal the one-GPU
solve of rank 0 BIT FOR BIT (fractions, T, rates, sweep count)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rabacus_b200 import configs  # noqa: E402
from rabacus_b200.plan import FIELDS, Plan  # noqa: E402


def stage(plan, cfg):
    plan.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"],
                     cfg["z"], 0.0)
    plan.upload()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_solves = int(sys.argv[sys.argv.index("--solves") + 1]) if "--solves" in sys.argv else 24
    Nl, Nnu, Nmu = 1024, 128, 64
    ok, sweeps_total, t0 = True, 0, time.perf_counter()
    for teq in (False, True):
        cfg = configs.c4_sphere_bgnd(Nl, Nnu, Nmu, teq, nH0=3.0)
        single = Plan(0, 1, Nl, Nnu, Nmu, find_Teq=teq, device=local)
        stage(single, cfg)
        ref_it = single.solve()["iters"]
        ref = single.get_problem(0)
        single.close()
        plan = Plan(0, 1, Nl, Nnu, Nmu, find_Teq=teq, device=local)
        stage(plan, cfg)
        plan.connect_peers(rank, world)
        for i in range(n_solves // 2):
            plan.upload()
            it = plan.solve_cell_sharded(rank, world)
            got = plan.get_problem(0)
            same = it == ref_it and all(np.array_equal(got[k], ref[k]) for k in FIELDS)
            ok = ok and same
            sweeps_total += it
            if not same:
                print("rank %d: solve %d (find_Teq=%s) differs from the one-GPU solve" % (rank, i, teq))
        dist.barrier()
        plan.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("soak: %d ranks, %d sharded solves, %d sweeps through the peer exchange in %.1f s: %s"
              % (world, n_solves, sweeps_total, time.perf_counter() - t0,
                 "every solve bit-identical to one GPU" if flag.item() else "MISMATCH"))
        if flag.item():
            print("soak OK")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
