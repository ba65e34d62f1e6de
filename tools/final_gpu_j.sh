#!/bin/bash
# One B200: full ncu captures of the final kernels (fused k = 4, product, wide Gram, wide s-step update).
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sell_tma -s 2 -c 1 -o gpurun_out/fj_sell_k4 -f \
    python tools/profile_target.py --what mpk --k 4 --reps 3 > gpurun_out/fj_sell_k4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sell_tma -s 2 -c 1 -o gpurun_out/fj_sell_spmv -f \
    python tools/profile_target.py --what spmv --reps 3 > gpurun_out/fj_sell_spmv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gram_wide|scg_update_wide|sell_tma_kernel<2' -s 3 -c 3 -o gpurun_out/fj_cg -f \
    python tools/profile_cg.py --its 8 > gpurun_out/fj_cg.log 2>&1
