#!/bin/bash
# heis_pipe v13 (own spins prefetched into registers, own ring = output staging): parity, then the 512^3 bench
out=${1:-r02h4}
mkdir -p gpurun_out/$out
timeout 900 python -m pytest tests/test_gpu_pipe.py -x -q -k "pipe_kernel" > gpurun_out/$out/pytest.log 2>&1; tail -4 gpurun_out/$out/pytest.log
bash profiles/r02/sweep.sh $out heis3d_512 30 "heis_pipe=-1" "heis_pipe_pub=4,heis_pipe_lead=48" "heis_pipe_pub=8,heis_pipe_lead=96" "heis_pipe_stages=6,heis_pipe_own=2" "heis_pipe_stages=5,heis_pipe_own=3" "heis_pipe_vec=2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heis_pipe -s 2 -c 1 -o gpurun_out/$out/heis_pipe -f python profiles/prof_run.py heis3d_512 3 > gpurun_out/$out/ncu.log 2>&1; tail -1 gpurun_out/$out/ncu.log
