#!/bin/bash
# Heisenberg-only GPU visit: parity tests, bench of heis3d_512, ncu of the step kernel.  bash profiles/gpu_heis.sh <tag>
tag=${1:-heis}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "heis or Heis or fused or slab" > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -15 $out/pytest.log
timeout 300 python bench.py --workload heis3d_512 --no-also --no-cpu --e2e-steps 0 --steps 20 > $out/bench_heis.json 2> $out/bench_heis.err; cat $out/bench_heis.json; tail -3 $out/bench_heis.err
for v in $FUSED_VARIANTS; do
  VEGAS_TUNE="$v" timeout 300 python bench.py --workload heis3d_512 --no-also --no-cpu --e2e-steps 0 --steps 20 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$v', d['value'], d['ms_per_step'], d['roofline']['frac'])
    else: print(l.rstrip())"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:heis_ -s 3 -c 2 -f -o $out/heis \
  python profiles/prof_run.py heis3d_512 4 > $out/ncu_heis.log 2>&1; tail -3 $out/ncu_heis.log
