#!/bin/bash
# round 2, first GPU contact of heis_pipe: parity tests, then heis3d_512 old (wave) vs new (pipe) with ring variations
mkdir -p gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q > gpurun_out/r02a/pytest_pipe.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a/pytest_pipe.log
tail -5 gpurun_out/r02a/pytest_pipe.log
B="timeout 300 python bench.py --workload heis3d_512 --steps 20 --warmup 3 --no-also --no-cpu --e2e-steps 0"
VEGAS_TUNE=heis_pipe=0 $B > gpurun_out/r02a/bench_wave.json 2> gpurun_out/r02a/bench_wave.err
$B > gpurun_out/r02a/bench_pipe.json 2> gpurun_out/r02a/bench_pipe.err
VEGAS_TUNE=heis_pipe_stages=5,heis_pipe_own=3 $B > gpurun_out/r02a/bench_pipe_5_3.json 2> gpurun_out/r02a/bench_pipe_5_3.err
VEGAS_TUNE=heis_pipe_stages=4,heis_pipe_own=2 $B > gpurun_out/r02a/bench_pipe_4_2.json 2> gpurun_out/r02a/bench_pipe_4_2.err
VEGAS_TUNE=heis_pipe_stages=7,heis_pipe_own=2 $B > gpurun_out/r02a/bench_pipe_7_2.json 2> gpurun_out/r02a/bench_pipe_7_2.err
for f in gpurun_out/r02a/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["kernel_family"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"])
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/r02a/*.err
