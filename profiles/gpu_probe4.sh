#!/bin/bash
out=gpurun_out/r01v; mkdir -p $out
python profiles/fcc_probe.py probe > $out/a.txt 2>&1
for poll in 0 2 20; do
  VEGAS_BENCH_POLL_MS=$poll python bench.py --workload heis_fcc_384 --no-also --no-cpu --e2e-steps 0 --steps 10 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('poll $poll', d['ms_per_step'], d['clocks'])" >> $out/a.txt 2>&1
done
for poll in 0 2 20; do
  VEGAS_BENCH_POLL_MS=$poll python bench.py --workload heis3d_512 --no-also --no-cpu --e2e-steps 0 --steps 100 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('heis512 poll $poll', d['ms_per_step'], d['clocks'])" >> $out/a.txt 2>&1
done
for poll in 0 2 20; do
  VEGAS_BENCH_POLL_MS=$poll python bench.py --workload ising3d_1024 --no-also --no-cpu --e2e-steps 0 --steps 100 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ising poll $poll', d['ms_per_step'], d['clocks'])" >> $out/a.txt 2>&1
done
cat $out/a.txt
