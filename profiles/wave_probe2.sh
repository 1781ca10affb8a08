#!/bin/bash
out=gpurun_out/${1:-wave2}; mkdir -p $out
for v in heis_wave_c=0 heis_wave_c=4 heis_wave_c=8 heis_wave_c=16; do
  VEGAS_TUNE="$v" timeout 300 ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:heis_stencil -c 1200 --csv --log-file $out/$v.csv python profiles/prof_run.py heis3d_512 2 > /dev/null 2>&1
  python - "$out/$v.csv" "$v" <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
tot={}
for r in rows: tot[r[-3]]=tot.get(r[-3],0.0)+float(r[-1].replace(',',''))
n=len(rows)//4
print('  ncu(no cache flush)', sys.argv[2], 'launches', n, {k:round(v/1e9,3) for k,v in tot.items()}, '(2 steps; GB / s)')
PY
done
