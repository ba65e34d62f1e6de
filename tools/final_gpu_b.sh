#!/bin/bash
# Investigation call: explicit-column fused kernel on the unstructured operator (variants, stage timing, one full ncu
# capture per variant) and the per-kernel decomposition of classical / s-step CG.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "## $1 at +$(( $(date +%s) - T0 )) s" | tee -a gpurun_out/fb_progress.log; }
stamp "pytest new variants"
timeout 600 python -m pytest tests/test_sell_gpu.py -x -q -k "spmv_and_mpk_bitwise" > gpurun_out/fb_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/fb_progress.log
stamp "C4 variants k=4"
timeout 600 python tools/profile_c4.py --m 200 --k 4 --variants 0,16,17 --timing > gpurun_out/fb_c4_k4.txt 2>&1
stamp "C4 variants k=8, k=2"
timeout 600 python tools/profile_c4.py --m 200 --k 8 --variants 0,16 > gpurun_out/fb_c4_k8.txt 2>&1
stamp "C4 ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sell_tma -c 2 -o gpurun_out/fb_c4_sell -f \
    python tools/profile_c4.py --m 200 --k 4 --variants 0,16 --ncu > gpurun_out/fb_c4_ncu.log 2>&1
stamp "CG timing"
timeout 600 python tools/profile_cg.py --its 32 > gpurun_out/fb_cg.txt 2>&1
stamp "CG launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/fb_cg_launches.csv \
    python tools/profile_cg.py --its 16 > gpurun_out/fb_cg_ncu.log 2>&1
stamp "done"
