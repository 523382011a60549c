#!/bin/bash
# launch shapes of the column kernel: one rank's share on one GPU (strides 1, -2, -4, -8)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for k in 11 12 13 9; do
echo "RB_COLS_VARIANT=$k" | tee -a gpurun_out/r02z_cols_variants.log
RB_COLS_VARIANT=$k timeout 300 python scratch/rank_share2.py 2>&1 | grep "stride -[248] rank 0\|stride 1 rank 0" | tee -a gpurun_out/r02z_cols_variants.log
done
