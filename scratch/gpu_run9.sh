#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python scratch/dbg_inputs.py
timeout 900 python -m pytest tests/test_gpu_gauss_seidel.py tests/test_gpu_teq_memo.py -m gpu -q -x > gpurun_out/r02g_pytest_gs.log 2>&1
echo "pytest gs rc=$?"; tail -25 gpurun_out/r02g_pytest_gs.log
timeout 900 python scratch/sanitize_small.py > gpurun_out/r02g_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -5 gpurun_out/r02g_sanitize_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sphere_columns_smem' -s 2 -c 1 -o gpurun_out/r02g_c5cols python scratch/c5_stages.py 1024 --sweeps 3 > gpurun_out/r02g_c5cols.log 2>&1
echo "ncu rc=$?"
