#!/bin/bash
# heis_pipe: DRAM traffic (ncu, one step) and step time (bench, 30 steps) by publication interval / lead
out=gpurun_out/r02h7; mkdir -p $out
for t in "heis_pipe_pub=4,heis_pipe_lead=32" "heis_pipe_pub=4,heis_pipe_lead=48" "heis_pipe_pub=6,heis_pipe_lead=40" "heis_pipe_pub=6,heis_pipe_lead=48" "heis_pipe_pub=8,heis_pipe_lead=48" "heis_pipe_pub=8,heis_pipe_lead=64"; do
  VEGAS_TUNE=$t timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:heis_pipe -s 2 -c 1 --csv python profiles/prof_run.py heis3d_512 3 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{printf "%s=%s%s ", $(NF-2), $NF, $(NF-1)}' | tr -d '"'
  VEGAS_TUNE=$t timeout 300 python bench.py --workload heis3d_512 --steps 30 --warmup 3 --no-also --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' | $t  %.4f ms'%d['ms_per_step'])"
done
