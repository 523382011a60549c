"""C5 (BASELINE.json configs[4]): parameter sweep of 8192 independent SphereBgnd solves
(32 nH0 x 16 R x 16 z; Nl=512, Nnu=128, Nmu=32), problems dealt (shuffled) over the
ranks (--block: contiguous blocks), no communication.  Prints one JSON line (rank 0): whole-job evals/s over full solves."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rabacus_b200 import configs
from rabacus_b200.plan import Plan, shard_problems, deal_problems

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
find_Teq = "--find_Teq" in sys.argv
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.perf_counter()
probs = configs.c5_sweep(find_Teq=find_Teq)
prank, pworld = rank, world
if "--as-rank" in sys.argv:   # one rank's share of an N-GPU run, alone on one GPU: --as-rank r/N
    prank, pworld = (int(v) for v in sys.argv[sys.argv.index("--as-rank") + 1].split("/"))
if "--block" in sys.argv:
    lo, hi = shard_problems(len(probs), prank, pworld)
    mine = probs[lo:hi]
else:
    mine = [probs[i] for i in deal_problems(len(probs), prank, pworld)]
t_build = time.perf_counter() - t0
plan = Plan(0, len(mine), 512, 128, 32, find_Teq=find_Teq, tol=1e-4, device=local)
t0 = time.perf_counter()
for b, c in enumerate(mine):
    plan.set_problem(b, c["Edges"], c["nH"], c["nHe"], c["T"], c["E_eV"], c["shape"], c["z"])
t_stage = time.perf_counter() - t0
def solve():
    plan.upload()
    return plan.solve()
solve()  # warm
torch.cuda.synchronize()
if dist: dist.barrier()
t0 = time.perf_counter()
info = solve()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ev = torch.tensor([float(info["evals"]), dt, info["ms_device"] * 1e-3], dtype=torch.float64, device="cuda")
if dist:
    evs = [torch.zeros_like(ev) for _ in range(world)]
    dist.all_gather(evs, ev)
else:
    evs = [ev]
if rank == 0:
    tot = sum(float(e[0]) for e in evs); tmax = max(float(e[1]) for e in evs)
    print(json.dumps({"workload": "C5: 8192 SphereBgnd solves 512x128x32%s" % (" find_Teq" if find_Teq else ""),
                      "n_gpus": world, "problems_per_rank": len(mine), "evals": tot, "wall_s": tmax,
                      "evals_per_s": tot / tmax, "max_sweeps": info["iters"],
                      "stage_ms_rank0": {k: round(info[k], 1) for k in ("ms_columns", "ms_nu", "ms_cell")},
                      "host_build_s": round(t_build, 2), "host_stage_s": round(t_stage, 2),
                      "per_rank_s": [round(float(e[1]), 3) for e in evs],
                      "dev_mem_used_gb_rank0": round((torch.cuda.mem_get_info()[1]
                                                      - torch.cuda.mem_get_info()[0]) / 2**30, 2)}))
if dist:
    dist.barrier(); dist.destroy_process_group()
