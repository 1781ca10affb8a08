#!/bin/bash
# Final GPU-box visit of round 2: parity tests, smoke, bench (ours + reference arm), ncu launch list and the `--set full`
# captures of the dominant kernels.   usage (under gpurun): bash profiles/r02/z_final.sh [tag]
tag=${1:-r02z}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --durations=8 > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
timeout 180 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 > $out/bench_under_ncu.log 2>&1
N="timeout 600 ncu --set full --clock-control none --import-source on"
$N -k regex:heis_pipe -s 2 -c 1 -o $out/heis_pipe -f python profiles/prof_run.py heis3d_512 3 > $out/ncu_heis_pipe.log 2>&1
# (gpurun brings back at most 64 MiB: the other captures are made by z_final2.sh in a call of their own)
tail -6 $out/pytest_gpu.log; tail -2 $out/smoke.log; cat $out/bench.json | cut -c1-600; tail -3 $out/bench.err; ls -la $out
