#!/bin/bash
# Per-kernel durations of the C5 step for each experiment build of libfluxgnn (see build.py --suffix).
# usage: scripts/fft_variants.sh "" _v1 _v2 ...   -> gpurun_out/variants.txt
out=gpurun_out/variants.txt; : > $out
for sfx in "$@"; do
  [ "$sfx" = base ] && sfx=""
  lib=$PWD/gnn_plasma_flux_b200/libfluxgnn${sfx}.so
  echo "== variant '${sfx}'" >> $out
  FLUXGNN_LIB=$lib python bench.py --workload c5 --steps 20 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))" >> $out
  FLUXGNN_LIB=$lib ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"poisson_fft|baseline_fv" -s 8 -c 4 --csv python bench.py --workload c5 --steps 3 --warmup 2 2>/dev/null | python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows: d.setdefault((r[0],r[4][:50]),{})[r[-3]]=r[-1]
for k,v in d.items(): print(k[1], v)
" >> $out
done
cat $out
