#!/bin/bash
# GPU run 1: parity tests, then the bench line (both arms), on one B200
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r02_gpu.txt 2>&1
nproc >> gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref_n1.json 2> gpurun_out/r02_ref_n1.err
echo "ref rc=$?"
cat gpurun_out/r02_ref_n1.json | head -c 600
