#!/bin/bash
# heis_pipe v4: progress published every K planes
mkdir -p gpurun_out/r02e
timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q > gpurun_out/r02e/pytest_pipe.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e/pytest_pipe.log
tail -3 gpurun_out/r02e/pytest_pipe.log
B="timeout 300 python bench.py --workload heis3d_512 --steps 20 --warmup 3 --no-also --no-cpu --e2e-steps 0"
for t in "heis_pipe_vec=4,heis_pipe_pub=1" "heis_pipe_vec=4,heis_pipe_pub=2" "heis_pipe_vec=4,heis_pipe_pub=4" "heis_pipe_vec=4,heis_pipe_pub=8" "heis_pipe_vec=2,heis_pipe_pub=2" "heis_pipe_vec=2,heis_pipe_pub=4" "heis_pipe_vec=2,heis_pipe_pub=8" "heis_pipe_vec=2,heis_pipe_pub=4,heis_pipe_lead=24" "heis_pipe_vec=4,heis_pipe_pub=4,heis_pipe_stages=5,heis_pipe_own=3"; do
  VEGAS_TUNE=$t $B > gpurun_out/r02e/bench_$t.json 2> gpurun_out/r02e/bench_$t.err
  python - "gpurun_out/r02e/bench_$t.json" "$t" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], d["kernel_family"], "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[2], "ERR", e)
PY
done
VEGAS_TUNE=heis_pipe_vec=4,heis_pipe_pub=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:heis_pipe -s 2 -c 1 -o gpurun_out/r02e/heis_pipe_v4 -f \
    python profiles/prof_run.py heis3d_512 3 > gpurun_out/r02e/ncu.log 2>&1
tail -2 gpurun_out/r02e/ncu.log
