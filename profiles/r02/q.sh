#!/bin/bash
mkdir -p gpurun_out/r02q
timeout 900 python -m pytest tests/test_gpu_pipe.py -x -q -k "pipe_kernel" 2>&1 | tail -2
bash profiles/r02/sweep.sh r02q heis3d_512 20 "heis_pipe=-1" "heis_pipe_lead=12" "heis_pipe_lead=24" "heis_pipe_lead=32" "heis_pipe_pub=2,heis_pipe_lead=24" "heis_pipe_pub=4,heis_pipe_lead=32" "heis_pipe_lead=16,heis_pipe_stages=5,heis_pipe_own=4"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heis_pipe -s 2 -c 1 -o gpurun_out/r02q/heis_pipe_v12 -f \
    python profiles/prof_run.py heis3d_512 3 > gpurun_out/r02q/ncu.log 2>&1
tail -2 gpurun_out/r02q/ncu.log
