#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export RB_PEER_TIMEOUT_MS=60000
which compute-sanitizer || export PATH=/usr/local/cuda/bin:$PATH
for tool in memcheck racecheck synccheck; do
  t0=$SECONDS
  timeout 1300 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python scratch/sanitize_small.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$? wall $((SECONDS-t0)) s"; echo "wall $((SECONDS-t0)) s" >> gpurun_out/r02_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ALL CASES DONE" gpurun_out/r02_sanitizer_$tool.log | tail -4
done
