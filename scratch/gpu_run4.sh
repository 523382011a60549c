#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_adversarial.txt
RB_PARITY_REPORT=$PWD/gpurun_out/r02_parity_adversarial.txt timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_far.json 2> gpurun_out/r02_bench_far.err
echo "bench rc=$?"; tail -5 gpurun_out/r02_bench_far.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_far.json'))
print("ms_per_step",d["ms_per_step"],"value %.3e"%d["value"], d["roofline"]["kernels"], d["roofline"]["ms_cell_solve"], d["roofline"]["ms_finish"])
print("solve",d["solve"],"c4teq",d["c4_find_Teq"]["ms_per_step"],d["c4_find_Teq"]["stages_ms"])
for k in ("c5","c5_find_Teq"): print(k,{a:d[k][a] for a in ("wall_s","staging_s","solve_s","sweeps_max","first_sweep_stage_ms")})
PY
