#!/bin/bash
# Final GPU-box visit of the round: parity tests, smoke, bench (ours + reference arm), resident-kernel probe, ncu launch list.
# usage (under gpurun): bash profiles/gpu_final.sh <tag>
tag=${1:-r01z}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 420 python -m pytest tests -m gpu -q --timeout 150 --durations=8 > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
timeout 60 python profiles/resident_probe.py > $out/resident_probe.txt 2>&1
timeout 420 python bench.py > $out/bench.json 2> $out/bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 > $out/bench_under_ncu.log 2>&1
tail -15 $out/pytest_gpu.log; tail -2 $out/smoke.log; cat $out/resident_probe.txt; cat $out/bench.json; tail -3 $out/bench.err
