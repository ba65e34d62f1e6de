# ncu --set full of the fused kernel on the slab operators of rank 0 / rank 1 of 2 (ghost rows above / below), one GPU
for r in 0 1; do
timeout 900 ncu --set full --clock-control none -k regex:sell_tma -s 3 -c 1 -o gpurun_out/r02_slab_full_r$r -f python tools/slab_probe.py --world 2 --rank $r --once > gpurun_out/r02_slab_full_r$r.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
