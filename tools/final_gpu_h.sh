#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python tools/slab_probe.py --world 2 --rank 1 --sweep > gpurun_out/fh_slab_w2_r1_sweep.txt 2>&1
for l2 in 60 70 75; do
timeout 200 python tools/slab_probe.py --world 2 --rank 1 --set wave_l2_pct=$l2 > gpurun_out/fh_slab_w2_r1_l2_$l2.txt 2>&1
done
timeout 300 python tools/slab_probe.py --world 2 --rank 0 --set wave_l2_pct=75 > gpurun_out/fh_slab_w2_r0_l2_75.txt 2>&1
