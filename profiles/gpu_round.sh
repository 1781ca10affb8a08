#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full of the two stencil kernels.
# usage (under gpurun): bash profiles/gpu_round.sh <tag>
tag=${1:-r01}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 > $out/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ising_msc -s 4 -c 2 -f -o $out/ising_msc \
  python profiles/prof_run.py ising3d_1024 4 > $out/ncu_ising.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:heis_stencil -s 4 -c 2 -f -o $out/heis_stencil \
  python profiles/prof_run.py heis3d_512 4 > $out/ncu_heis.log 2>&1
tail -3 $out/pytest_gpu.log; cat $out/bench.json
