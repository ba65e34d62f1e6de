#!/bin/bash
# Two B200s: multi-GPU parity check (strategies, push halo, unstructured operator, CG, rank-local reductions) and the bench
# line with parity, strong sub-record and the config-5 CG record (small grid: the code path bench.py --gpus 8 takes).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "## $1 at +$(( $(date +%s) - T0 )) s" | tee -a gpurun_out/fd_progress.log; }
stamp "dist_check"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/dist_check.py > gpurun_out/fd_dist_check.txt 2>&1
echo "dist_check rc=$?" >> gpurun_out/fd_progress.log
stamp "bench 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 3 --c5 --c5-grid 256 > gpurun_out/fd_bench_2gpu.json 2> gpurun_out/fd_bench_2gpu.err
echo "bench rc=$?" >> gpurun_out/fd_progress.log
stamp "done"
