#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for r in 1 2 1 2; do RB_NU_RU=$r timeout 300 python scratch/c5_stages.py 2048; done
