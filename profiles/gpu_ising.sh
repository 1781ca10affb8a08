#!/bin/bash
# Ising iteration: parity tests, bench, optional ncu.   bash profiles/gpu_ising.sh <tag> [ncu]
tag=${1:-ising}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "ising or Ising or slab or machine or host or onsager or exact" 2>&1 | tail -4
for w in ising3d_1024 ising2d_8192; do
timeout 300 python bench.py --workload $w --no-also --no-cpu --e2e-steps 0 --steps 50 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$w', '%.4g attempts/s' % d['value'], '%.4f ms/step' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'])
    else: print(l.rstrip())"
done
if [ "$2" = "ncu" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:ising_msc -s 4 -c 2 -f -o $out/ising_msc \
    python profiles/prof_run.py ising3d_1024 4 > $out/ncu.log 2>&1; tail -2 $out/ncu.log
fi
