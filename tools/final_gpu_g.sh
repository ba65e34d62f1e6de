#!/bin/bash
# One B200: counters of the fused kernel on the two slab operators of a 2-rank weak split, then the round-end validation
# (the -m gpu suite, smoke, the bench line, the launch list) on the final code.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "## $1 at +$(( $(date +%s) - T0 )) s" | tee -a gpurun_out/fg_progress.log; }
stamp "slab counters"
for r in 0 1; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:sell_tma -s 3 -c 1 --csv --log-file gpurun_out/fg_slab_r$r.csv python tools/slab_probe.py --world 2 --rank $r --once > gpurun_out/fg_slab_r$r.log 2>&1
done
stamp "pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/fg_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/fg_progress.log
stamp "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fg_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/fg_progress.log
stamp "bench"
timeout 600 python bench.py > gpurun_out/fg_bench_1gpu.json 2> gpurun_out/fg_bench_1gpu.err
echo "bench rc=$?" >> gpurun_out/fg_progress.log
stamp "launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fg_launch_list.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/fg_launch_list.log 2>&1
stamp "done"
