#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python scratch/rank_share2.py
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multigpu/check_cell_sharding.py > gpurun_out/r02l_check_n2.log 2>&1; echo "check rc=$?"; grep -v "^\*\*\*\|NCCL version\|OMP_NUM" gpurun_out/r02l_check_n2.log | tail -16
timeout 600 $TR --master-port 29512 tests/multigpu/soak_peer.py --solves 8 > gpurun_out/r02l_soak_n2.log 2>&1; echo "soak rc=$?"; tail -3 gpurun_out/r02l_soak_n2.log
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r02l_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench_n2.json'))
print("N=2 ms_per_step",d["ms_per_step"],"value %.3e"%d["value"], d["roofline"]["kernels"], d["roofline"]["ms_cell_solve"], d["roofline"]["ms_finish"])
print("parity_vs_n1", d["parity_vs_n1"])
PY
