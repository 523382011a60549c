#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z3_bench_n1.json 2> gpurun_out/r02z3_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r02z3_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02z3_bench_n1.json')); print('ms_per_step', d['ms_per_step'], 'value %.3e e2e %.3e' % (d['value'], d['e2e']['value']), d['roofline']['kernels'], 'c5', d['c5']['wall_s'], d['c5_find_Teq']['wall_s'], d['clocks'])"
