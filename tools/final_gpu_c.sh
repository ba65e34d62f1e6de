#!/bin/bash
# Third call: full -m gpu suite on the changed kernels (wide Gram / s-step update, 256-bit block loads, explicit-column
# default), CG decomposition, block CSR at two sizes, window / chunk / hint sweep of the explicit-column fused kernel,
# C4 at three sizes incl. its stated 49.8 M rows.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "## $1 at +$(( $(date +%s) - T0 )) s" | tee -a gpurun_out/fc_progress.log; }
stamp "pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/fc_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/fc_progress.log
stamp "CG timing"
timeout 600 python tools/profile_cg.py --its 32 > gpurun_out/fc_cg.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/fc_cg_launches.csv \
    python tools/profile_cg.py --its 16 > gpurun_out/fc_cg_ncu.log 2>&1
stamp "C1 block CSR"
timeout 600 python tools/bench_configs.py --only c1 > gpurun_out/fc_configs_c1.txt 2>&1
timeout 900 python tools/bench_configs.py --only c1 --c1 72 > gpurun_out/fc_configs_c1_72.txt 2>&1
stamp "C4 sweep k=4"
timeout 900 python tools/profile_c4.py --m 200 --k 4 --variants 0 --l2 0,40,55,70 --chunks 0,1,2 --flags -1,11 > gpurun_out/fc_c4_sweep_k4.txt 2>&1
stamp "C4 k=8, 1 M and 8.1 M rows"
timeout 600 python tools/profile_c4.py --m 100 --k 8 --variants 0,8 > gpurun_out/fc_c4_m100_k8.txt 2>&1
timeout 600 python tools/profile_c4.py --m 200 --k 8 --variants 0 --l2 0,55,70 > gpurun_out/fc_c4_m200_k8.txt 2>&1
stamp "C4 at 49.8 M rows"
timeout 900 python tools/bench_c4.py --m 367 --reps 4 > gpurun_out/fc_c4_m367.txt 2>&1
stamp "done"
