#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/r02w_smoke.log
python scratch/zones_once.py
timeout 600 ncu --set full --clock-control none -k regex:'k_zone_pce|k_zone_pcte|k_kchem' -s 3 -c 3 -o gpurun_out/r02w_zones python scratch/zones_once.py > gpurun_out/r02w_zones_ncu.log 2>&1; echo "ncu rc=$?"
