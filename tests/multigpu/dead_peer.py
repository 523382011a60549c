"""A rank that stops sweeping must surface as RB_ERR_PEER_TIMEOUT on its peers, not as a hang
(run under torchrun with 2 ranks, RB_PEER_TIMEOUT_MS=400):

rank 1 takes part in three sharded sweeps and then simply stops; rank 0 runs the whole
sharded solve.  Its fourth barrier never completes: the finish kernel gives up after the
timeout, poisons the flags and the solve returns error 12."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("RB_PEER_TIMEOUT_MS", "400")
from rabacus_b200 import configs  # noqa: E402
from rabacus_b200._lib import RabacusB200Error  # noqa: E402
from rabacus_b200.plan import Plan  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = configs.c4_sphere_bgnd(256, 64, 8, False, nH0=3.0)
    plan = Plan(0, 1, 256, 64, 8, device=local)
    plan.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"], 3.0)
    plan.upload()
    plan.connect_peers(rank, world)
    dist.barrier()
    code, dt = 0, 0.0
    if rank == 0:
        t0 = time.perf_counter()
        try:
            plan.solve_cell_sharded(rank, world)
        except RabacusB200Error as e:
            code = e.code
        dt = time.perf_counter() - t0
        print("rank 0: sharded solve with a dead peer returned error %d after %.2f s (%s)"
              % (code, dt, "RB_ERR_PEER_TIMEOUT" if code == 12 else "UNEXPECTED"))
    else:
        plan.thin()
        for _ in range(3):
            plan.sweep_sharded()
            plan.poll_sweep()
        time.sleep(6.0)   # "dead": alive as a process, absent from the barrier
    ok = torch.tensor([1 if (rank != 0 or (code == 12 and dt < 30.0)) else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0 and ok.item():
        print("dead-peer check OK")
    try:
        plan.close()
    except Exception:
        pass
    dist.destroy_process_group()
    sys.exit(0 if ok.item() else 1)


if __name__ == "__main__":
    main()
