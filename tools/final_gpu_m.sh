#!/bin/bash
# Validation after the window default moved to 92 % for pattern operators: suite, bench, DRAM traffic of the fused launch.
set -u
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/fm_bench_1gpu.json 2> gpurun_out/fm_bench_1gpu.err
echo "bench rc=$?" >> gpurun_out/fm_bench_1gpu.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:sell_tma -s 2 -c 2 --csv --log-file gpurun_out/fm_traffic.csv python tools/profile_target.py --what mpk --k 4 --reps 3 > gpurun_out/fm_traffic.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/fm_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/fm_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fm_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/fm_smoke.log
