#!/bin/bash
# L2-reuse experiment: DRAM bytes and time of one Heisenberg 512^3 step for several launch orders
out=gpurun_out/${1:-wave}; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q -k "wave or fused" 2>&1 | tail -2
for v in heis_wave_c=0 heis_wave_c=2 heis_wave_c=4 heis_wave_c=8 heis_wave_c=16 heis_fused=1; do
  VEGAS_TUNE="$v" timeout 300 python bench.py --workload heis3d_512 --no-also --no-cpu --e2e-steps 0 --steps 10 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$v', '%.4g attempts/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'launches', d['gpu_launches'])
    else: print(l.rstrip())"
  VEGAS_TUNE="$v" timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:heis_ -c 1200 --csv --log-file $out/$v.csv python profiles/prof_run.py heis3d_512 2 > /dev/null 2>&1
  python - "$out/$v.csv" "$v" <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
tot={}
for r in rows: tot[r[-3]]=tot.get(r[-3],0.0)+float(r[-1].replace(',',''))
units={r[-3]:r[-2] for r in rows}
n=len(rows)//3
print('  ncu', sys.argv[2], 'launches', n, {k:(round(v,3),units[k]) for k,v in tot.items()}, '(2 steps)')
PY
done
