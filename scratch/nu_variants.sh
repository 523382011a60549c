for v in 0 1 2 3; do RB_NU_VARIANT=$v python scratch/c5_variants.py c4 2>&1 | tail -1; done
python scratch/c5_variants.py c5 2>&1 | tail -1
