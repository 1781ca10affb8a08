#!/bin/bash
# heis_pipe: L2 eviction-priority hints on the TMA traffic -- parity, traffic (ncu) and time
out=gpurun_out/r02h8; mkdir -p $out
VEGAS_TUNE_TEST=1 timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q -k "pipe_kernel" > $out/pytest.log 2>&1; tail -2 $out/pytest.log
for t in "heis_pipe_l2=0" "heis_pipe_l2=1" "heis_pipe_l2=1,heis_pipe_pub=4,heis_pipe_lead=32" "heis_pipe_l2=1,heis_pipe_pub=8,heis_pipe_lead=64" "heis_pipe_l2=1,heis_pipe_pub=2,heis_pipe_lead=24"; do
  VEGAS_TUNE=$t timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:heis_pipe -s 2 -c 1 --csv python profiles/prof_run.py heis3d_512 3 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{printf "%s=%s%s ", $(NF-2), $NF, $(NF-1)}' | tr -d '"'
  VEGAS_TUNE=$t timeout 300 python bench.py --workload heis3d_512 --steps 30 --warmup 3 --no-also --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' | $t  %.4f ms'%d['ms_per_step'])"
done
