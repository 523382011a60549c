#!/bin/bash
# ncu source-level capture of one rank's share of an 8-GPU C4 sweep (alone on one GPU)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scratch/rank_share_prof.py 8 > gpurun_out/r02z_share8_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_far_moments|k_sphere_columns|k_columns_combine|k_nu_rates' -s 8 -c 4 -o gpurun_out/r02z_share8 python scratch/rank_share_prof.py 8 > gpurun_out/r02z_share8_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
