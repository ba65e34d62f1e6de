#!/bin/bash
# Two B200s: config 4 (RCM-ordered tet P1, 8.1 M rows) cut into two row blocks with a depth-8 halo.
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    tools/dist_c4.py --mesh 200 --k 8 > gpurun_out/fe_dist_c4.txt 2>&1
echo "rc=$?" >> gpurun_out/fe_dist_c4.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
    tools/dist_c4.py --mesh 100 --k 8 > gpurun_out/fe_dist_c4_m100.txt 2>&1
echo "rc=$?" >> gpurun_out/fe_dist_c4_m100.txt
