#!/bin/bash
# One B200: the slab operators of a multi-GPU split in isolation (no exchange): where the weak-scaling gap comes from.
set -u
mkdir -p gpurun_out
for wr in "2 0" "2 1" "8 3"; do
  set -- $wr
  timeout 300 python tools/slab_probe.py --world $1 --rank $2 --timing > gpurun_out/ff_slab_w$1_r$2.txt 2>&1
done
timeout 300 python tools/slab_probe.py --world 2 --rank 1 --set sell_geom=2 > gpurun_out/ff_slab_w2_r1_masked.txt 2>&1
timeout 300 python tools/slab_probe.py --world 8 --rank 3 --strong --timing > gpurun_out/ff_slab_w8_r3_strong.txt 2>&1
