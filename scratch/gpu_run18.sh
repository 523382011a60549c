#!/bin/bash
# round-2 final records on one GPU: bench line, reference arm, launch list, ncu --set full (default
# cache control = flushed between replays, and --cache-control none = as in the pipeline)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02y_bench_n1.json 2> gpurun_out/r02y_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02y_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02y_ref_n1.json 2> gpurun_out/r02y_ref_n1.err; echo "ref rc=$?"; tail -2 gpurun_out/r02y_ref_n1.err; cat gpurun_out/r02y_ref_n1.json | cut -c1-400
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02y_bench_n1.json'))
print("ms_per_step",d["ms_per_step"],"value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], d["roofline"]["kernels"], d["roofline"]["ms_cell_solve"], d["roofline"]["ms_finish"])
print("solve",d["solve"],"c4teq",d["c4_find_Teq"]["ms_per_step"],d["c4_find_Teq"]["stages_ms"], d["c4_find_Teq"].get("solve"))
for k in ("c5","c5_find_Teq"): print(k,{a:d[k].get(a) for a in ("wall_s","staging_s","solve_s","readback_s","plan_create_s","sweeps_max")})
print("zones", json.dumps(d["zones"])[:900]); print("cpu", json.dumps(d["cpu_baseline"])[:600]); print("clocks", d["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02y_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r02y_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_nu_rates|k_sphere_columns|k_far_moments|k_cell_solve' -c 8 -o gpurun_out/r02y_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r02y_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum -k regex:'k_nu_rates|k_sphere_columns|k_far_moments' -c 9 --csv --log-file gpurun_out/r02y_traffic_nocachectl.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r02y_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
