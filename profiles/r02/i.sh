#!/bin/bash
mkdir -p gpurun_out/r02i
timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q > gpurun_out/r02i/pytest.log 2>&1; tail -3 gpurun_out/r02i/pytest.log
bash profiles/r02/sweep.sh r02i heis3d_512 20 "heis_pipe_tstore=0,heis_pipe_vec=4,heis_pipe_pub=8,heis_pipe_lead=48" "heis_pipe_tstore=1,heis_pipe_vec=4,heis_pipe_pub=8,heis_pipe_lead=48" "heis_pipe_tstore=1,heis_pipe_vec=4,heis_pipe_pub=4,heis_pipe_lead=32" "heis_pipe_tstore=1,heis_pipe_vec=4,heis_pipe_pub=2,heis_pipe_lead=24"
VEGAS_TUNE=heis_pipe_tstore=0,heis_pipe_vec=4,heis_pipe_pub=8,heis_pipe_lead=48 timeout 900 ncu --set full --clock-control none --import-source on -k regex:heis_pipe -s 2 -c 1 -o gpurun_out/r02i/heis_pipe_v8 -f \
    python profiles/prof_run.py heis3d_512 3 > gpurun_out/r02i/ncu.log 2>&1
tail -2 gpurun_out/r02i/ncu.log
