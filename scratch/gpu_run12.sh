#!/bin/bash
# 2-GPU run: sharding check, soak, dead peer, bench N=2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multigpu/check_cell_sharding.py > gpurun_out/r02_check_cell_sharding_n2.log 2>&1; echo "check rc=$?"; tail -16 gpurun_out/r02_check_cell_sharding_n2.log
timeout 600 $TR --master-port 29512 tests/multigpu/soak_peer.py --solves 24 > gpurun_out/r02_soak_n2.log 2>&1; echo "soak rc=$?"; tail -4 gpurun_out/r02_soak_n2.log
timeout 120 $TR --master-port 29513 tests/multigpu/dead_peer.py > gpurun_out/r02_dead_peer_n2.log 2>&1; echo "dead rc=$?"; tail -4 gpurun_out/r02_dead_peer_n2.log
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print("N=2 ms_per_step",d["ms_per_step"],"value %.3e"%d["value"], d["roofline"]["kernels"], d["roofline"]["ms_cell_solve"], d["roofline"]["ms_finish"])
print("parity_vs_n1", d["parity_vs_n1"]); print("c4teq", d["c4_find_Teq"]["ms_per_step"], d["c4_find_Teq"]["stages_ms"])
for k in ("c5","c5_find_Teq"): print(k,{a:d[k].get(a) for a in ("wall_s","staging_s","solve_s","sweeps_max")})
PY
