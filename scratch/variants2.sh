python scratch/c5_variants.py host 2>&1 | tail -6
for v in 0 1 2 3 4 5 6; do RB_NU_VARIANT=$v python scratch/c5_variants.py c5 2>&1 | tail -1; done
for w in 1 2 8; do RB_COLS_WAVES=$w python scratch/c5_variants.py c5 2>&1 | tail -1; done
for v in 0 1 2 3 4 5; do RB_NU_VARIANT=$v python scratch/c5_variants.py c4 2>&1 | tail -1; done
