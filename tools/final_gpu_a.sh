#!/bin/bash
# Round-end validation on ONE B200 (run under gpurun): the -m gpu suite, the bench line and the reference arm, the launch
# list of the bench command, the block-CSR / C1 numbers, and L2 / DRAM evidence for the unstructured (C4) operator.
# Everything lands in gpurun_out/; nothing here is a bench value when it ran under ncu.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "## $1 at +$(( $(date +%s) - T0 )) s" | tee -a gpurun_out/fa_progress.log; }

stamp "pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/fa_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/fa_progress.log
tail -3 gpurun_out/fa_pytest.log

stamp "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fa_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/fa_progress.log

stamp "bench"
timeout 600 python bench.py > gpurun_out/fa_bench_1gpu.json 2> gpurun_out/fa_bench_1gpu.err
echo "bench rc=$?" >> gpurun_out/fa_progress.log
stamp "bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fa_bench_ref.json 2> gpurun_out/fa_bench_ref.err

stamp "launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fa_launch_list.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/fa_launch_list.log 2>&1

stamp "C1 (block CSR) numbers"
timeout 600 python tools/bench_configs.py --only c1 > gpurun_out/fa_configs_c1.txt 2>&1

stamp "C4 8.1 M rows: timings"
timeout 600 python tools/bench_c4.py --m 200 --reps 4 > gpurun_out/fa_c4_m200.txt 2>&1
stamp "C4 8.1 M rows: L2 / DRAM counters per kernel"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:'sell|spmv_stream|packed' -c 70 --csv --log-file gpurun_out/fa_c4_ncu.csv \
    python tools/bench_c4.py --m 200 --reps 2 > gpurun_out/fa_c4_ncu.log 2>&1
stamp "done"
