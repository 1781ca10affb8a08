#!/bin/bash
# persistent wave kernel: tests, then bench variants (each under its own timeout).  bash profiles/gpu_wave.sh <tag> "<variants>"
tag=${1:-wave}; out=gpurun_out/$tag; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q -k "wave_kernel" 2>&1 | tail -3
for v in $2; do
  VEGAS_TUNE="$v" timeout 120 python bench.py --workload heis3d_512 --no-also --no-cpu --e2e-steps 0 --steps 20 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$v', '%.4g attempts/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])
    else: print(l.rstrip()[:300])"
done
if [ -n "$3" ]; then
  VEGAS_TUNE="$3" timeout 300 ncu --set full --clock-control none --import-source on -k regex:heis_wave -s 1 -c 1 -f -o $out/wave python profiles/prof_run.py heis3d_512 3 > $out/ncu.log 2>&1; tail -2 $out/ncu.log
fi
