#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( python scratch/adv_debug.py slab_plane_1 contrast12 0 2
  python scratch/adv_debug.py slab_plane_2 contrast12 0 2
  python scratch/adv_debug.py slab_plane_1 sharp_front 0 2
  python scratch/adv_debug.py sphere_bgnd contrast12 1 40
  python scratch/adv_debug.py slab_bgnd contrast12 1 2 ) > gpurun_out/r02_adv_debug.txt 2>&1
cat gpurun_out/r02_adv_debug.txt
