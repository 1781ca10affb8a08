#!/bin/bash
# pipe-rate microbenchmark; heis_pipe v3 (first colour bounded ahead of the second): tests, lead / vec sweep
mkdir -p gpurun_out/r02d
./profiles/micro/pipe_rates > gpurun_out/r02d/pipe_rates.txt 2>&1; cat gpurun_out/r02d/pipe_rates.txt
timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q > gpurun_out/r02d/pytest_pipe.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d/pytest_pipe.log
tail -3 gpurun_out/r02d/pytest_pipe.log
B="timeout 300 python bench.py --workload heis3d_512 --steps 20 --warmup 3 --no-also --no-cpu --e2e-steps 0"
for t in "heis_pipe_vec=4" "heis_pipe_vec=2" "heis_pipe_vec=4,heis_pipe_lead=4" "heis_pipe_vec=4,heis_pipe_lead=16" "heis_pipe_vec=2,heis_pipe_lead=4" "heis_pipe_vec=2,heis_pipe_lead=16" "heis_pipe_vec=2,heis_pipe_stages=5,heis_pipe_own=2" "heis_pipe_vec=4,heis_pipe_stages=5,heis_pipe_own=2"; do
  VEGAS_TUNE=$t $B > gpurun_out/r02d/bench_$t.json 2> gpurun_out/r02d/bench_$t.err
  python - "gpurun_out/r02d/bench_$t.json" "$t" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], d["kernel_family"], "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[2], "ERR", e)
PY
done
