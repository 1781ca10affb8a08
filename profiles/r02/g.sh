#!/bin/bash
mkdir -p gpurun_out/r02g
timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q 2>&1 | tail -2
bash profiles/r02/sweep.sh r02g heis3d_512 20 "heis_pipe_vec=4,heis_pipe_pub=4" "heis_pipe_vec=4,heis_pipe_pub=8" "heis_pipe_vec=4,heis_pipe_pub=16" "heis_pipe_vec=2,heis_pipe_pub=8" "heis_pipe_vec=4,heis_pipe_pub=8,heis_pipe_stages=5,heis_pipe_own=4" "heis_pipe_vec=4,heis_pipe_pub=8,heis_pipe_lead=40" "heis_pipe=0"
VEGAS_TUNE=heis_pipe_vec=4,heis_pipe_pub=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:heis_pipe -s 2 -c 1 -o gpurun_out/r02g/heis_pipe_v6 -f \
    python profiles/prof_run.py heis3d_512 3 > gpurun_out/r02g/ncu.log 2>&1
tail -2 gpurun_out/r02g/ncu.log
