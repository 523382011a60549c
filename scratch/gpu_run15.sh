#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sphere_columns_smem|k_far_moments' -s 4 -c 2 -o gpurun_out/r02k_share8 python scratch/share8_once.py > gpurun_out/r02k_share8.log 2>&1
echo "ncu rc=$?"
