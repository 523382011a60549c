#!/bin/bash
# N-GPU bench line (N = number of GPUs of the box) + sharding check + soak at that N
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02z_bench_n$N.json 2> gpurun_out/r02z_bench_n$N.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/r02z_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02z_bench_n$N.json'))
print("N=$N ms_per_step",d["ms_per_step"],"value %.3e"%d["value"], d["roofline"]["kernels"], d["roofline"]["ms_cell_solve"], d["roofline"]["ms_finish"])
print("solve", d["solve"], "parity_vs_n1", d["parity_vs_n1"]); print("c4teq", d["c4_find_Teq"]["ms_per_step"], d["c4_find_Teq"]["stages_ms"])
for k in ("c5","c5_find_Teq"): print(k,{a:d[k].get(a) for a in ("wall_s","staging_s","solve_s","readback_s","sweeps_max")})
PY
timeout 600 $TR --master-port 29511 tests/multigpu/check_cell_sharding.py > gpurun_out/r02z_check_n$N.log 2>&1; echo "check rc=$?"; grep -v "^\*\*\*\|NCCL version\|OMP_NUM" gpurun_out/r02z_check_n$N.log | tail -5
timeout 600 $TR --master-port 29512 tests/multigpu/soak_peer.py --solves 24 > gpurun_out/r02z_soak_n$N.log 2>&1; echo "soak rc=$?"; tail -3 gpurun_out/r02z_soak_n$N.log
