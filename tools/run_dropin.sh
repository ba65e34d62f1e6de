#!/bin/bash
# Runs the reference's own 2xSpMV benchmark driver (mpk/2SpMV.cpp, unchanged) twice on the same generated
# FEM-like matrix: once linked against the reference CPU kernels, once against the GPU shim.
# Build first (development container): make -C oracle dropin
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/write_mtx.py --kind fem --m ${1:-14} gpurun_out/fem.mtx
echo "== reference CPU kernels (mpk/SpMV.cpp) =="
oracle/_ref/2spmv_cpu gpurun_out/fem.mtx
echo "== GPU shim (libnsk_spmvshim.so) =="
oracle/_ref/2spmv_gpu gpurun_out/fem.mtx
