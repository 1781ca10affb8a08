#!/bin/bash
mkdir -p gpurun_out/r02k
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_parity.py -x -q -k "pipe or heisenberg" > gpurun_out/r02k/pytest.log 2>&1; tail -3 gpurun_out/r02k/pytest.log
bash profiles/r02/sweep.sh r02k heis3d_512 20 "heis_pipe_vec=4,heis_pipe_pub=4,heis_pipe_lead=32" "heis_pipe_vec=4,heis_pipe_pub=4,heis_pipe_lead=24" "heis_pipe_vec=4,heis_pipe_pub=2,heis_pipe_lead=16" "heis_pipe_vec=4,heis_pipe_pub=2,heis_pipe_lead=24" "heis_pipe_vec=4,heis_pipe_pub=8,heis_pipe_lead=48" "heis_pipe_vec=4,heis_pipe_pub=4,heis_pipe_lead=32,heis_pipe_stages=5,heis_pipe_own=3" "heis_pipe_vec=4,heis_pipe_pub=4,heis_pipe_lead=32,heis_pipe_stages=6,heis_pipe_own=3" "heis_pipe_vec=2,heis_pipe_pub=4,heis_pipe_lead=32"
VEGAS_TUNE=heis_pipe_vec=4,heis_pipe_pub=4,heis_pipe_lead=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:heis_pipe -s 2 -c 1 -o gpurun_out/r02k/heis_pipe_v10 -f \
    python profiles/prof_run.py heis3d_512 3 > gpurun_out/r02k/ncu.log 2>&1
tail -2 gpurun_out/r02k/ncu.log
