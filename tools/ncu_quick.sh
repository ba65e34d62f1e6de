#!/bin/bash
# usage: tools/ncu_quick.sh <tag> <profile_target args...>   -- DRAM bytes / L2 hit rate / duration of every packed|mpk|spmv kernel launch
tag=$1; shift
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_bytes.sum \
    --clock-control none -k regex:'packed_kernel|mpk_|spmv_stream' -s 1 -c 2 --csv --log-file gpurun_out/ncuq_$tag.csv \
    python tools/profile_target.py "$@" > gpurun_out/ncuq_$tag.log 2>&1
python - "$tag" <<'PY'
import csv,sys
tag=sys.argv[1]
rows=[r for r in csv.reader(open(f"gpurun_out/ncuq_{tag}.csv")) if len(r)>10]
hdr=rows[0]
i_id,i_k,i_m,i_u,i_v=hdr.index("ID"),hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Unit"),hdr.index("Metric Value")
out={}
for r in rows[1:]:
    out.setdefault((r[i_id],r[i_k][:40]),[]).append(f"{r[i_m].split('__')[1][:28]}={r[i_v]}{r[i_u]}")
for k,v in out.items(): print(tag,k[0],k[1],' '.join(v))
PY
