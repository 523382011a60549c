#!/bin/bash
# sparse-launch variants of the column kernel (cell shards over 4-8 GPUs), one rank's share on one GPU
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for k in 1 0 2 3; do
echo "RB_COLS_SPARSE=$k"
RB_COLS_SPARSE=$k timeout 300 python scratch/rank_share2.py 2>&1 | grep "stride -[48] rank 0\|stride -2 rank 0" | tee -a gpurun_out/r02z_sparse_variants.log
done
