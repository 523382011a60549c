#!/bin/bash
# final records on one GPU after k_far_moments_split: bench line, launch list, smoke
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02z_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_bench_n1.json'))
print("ms_per_step",d["ms_per_step"],"value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], d["roofline"]["kernels"], d["roofline"]["ms_cell_solve"], d["roofline"]["ms_finish"])
print("solve",d["solve"],"c4teq",d["c4_find_Teq"]["ms_per_step"],d["c4_find_Teq"]["stages_ms"], d["c4_find_Teq"].get("solve"))
for k in ("c5","c5_find_Teq"): print(k,{a:d[k].get(a) for a in ("wall_s","staging_s","solve_s","readback_s","plan_create_s","sweeps_max")})
print("clocks", d["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r02z_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02z2_smoke.log
