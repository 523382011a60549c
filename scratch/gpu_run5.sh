#!/bin/bash
# round-2 baseline of the far-field build: GPU tests, bench line, launch list, ncu --set full of the hot kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_adversarial.txt
RB_PARITY_REPORT=$PWD/gpurun_out/r02_parity_adversarial.txt timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02c_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
echo "bench rc=$?"; tail -5 gpurun_out/r02c_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench_n1.json'))
print("ms_per_step",d["ms_per_step"],"value %.3e"%d["value"], d["roofline"]["kernels"], d["roofline"]["ms_cell_solve"], d["roofline"]["ms_finish"])
print("solve",d["solve"],"c4teq",d["c4_find_Teq"]["ms_per_step"],d["c4_find_Teq"]["stages_ms"])
for k in ("c5","c5_find_Teq"): print(k,{a:d[k].get(a) for a in ("wall_s","staging_s","solve_s","sweeps_max","first_sweep_stage_ms")})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r02c_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_nu_rates|k_sphere_columns|k_far_moments|k_cell_solve' -c 8 -o gpurun_out/r02c_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r02c_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
