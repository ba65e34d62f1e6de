#!/bin/bash
# Last validation of the round on the final code: the -m gpu suite, smoke, the bench line.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/fk_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/fk_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fk_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/fk_smoke.log
timeout 600 python bench.py > gpurun_out/fk_bench_1gpu.json 2> gpurun_out/fk_bench_1gpu.err
echo "bench rc=$?" >> gpurun_out/fk_bench_1gpu.err
