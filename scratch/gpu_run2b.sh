#!/bin/bash
# 2 GPUs: sharded-path correctness (peer memory + NCCL) and the N=2 bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu/check_cell_sharding.py > gpurun_out/r02_check2.log 2>&1
echo "check rc=$?"; tail -30 gpurun_out/r02_check2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r02_bench_n2.err; head -c 400 gpurun_out/r02_bench_n2.json
