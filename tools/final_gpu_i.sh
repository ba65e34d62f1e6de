#!/bin/bash
# Four B200s: ranks with TWO neighbours (halo push to both sides): parity check and the bench line.
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 \
    tools/dist_check.py > gpurun_out/fi_dist_check4.txt 2>&1
echo "dist_check rc=$?" >> gpurun_out/fi_dist_check4.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/fi_bench_4gpu.json 2> gpurun_out/fi_bench_4gpu.err
echo "bench rc=$?" >> gpurun_out/fi_bench_4gpu.err
