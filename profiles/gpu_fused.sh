#!/bin/bash
# fused-kernel iteration: tests, bench (fused on/off), optional ncu.  bash profiles/gpu_fused.sh <tag> [ncu]
tag=${1:-fused}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or wave" 2>&1 | tail -4
for v in heis_fused=1 $EXTRA_VARIANTS; do
  VEGAS_TUNE="$v" timeout 300 python bench.py --workload heis3d_512 --no-also --no-cpu --e2e-steps 0 --steps 20 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$v', '%.4g attempts/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])
    else: print(l.rstrip())"
done
if [ "$2" = "ncu" ]; then
  VEGAS_TUNE="heis_fused=1" timeout 400 ncu --set full --clock-control none --import-source on -k regex:heis_fused -s 1 -c 1 -f -o $out/heis \
    python profiles/prof_run.py heis3d_512 3 > $out/ncu_heis.log 2>&1; tail -2 $out/ncu_heis.log
fi
