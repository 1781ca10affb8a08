#!/bin/bash
mkdir -p gpurun_out/r02p
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02p/pytest_gpu.log 2>&1; tail -5 gpurun_out/r02p/pytest_gpu.log
bash profiles/r02/sweep.sh r02p heis3d_512 20 "heis_pipe=-1" "heis_pipe_pub=2,heis_pipe_lead=24" "heis_pipe_pub=8,heis_pipe_lead=48" "heis_pipe=0" "heis_pipe=0,heis_wave=0"
bash profiles/r02/sweep.sh r02p heis_fcc_384 10 "basis_pipe=0" "basis_pipe=1"
bash profiles/r02/sweep.sh r02p ising3d_1024 30 "resident_max=8192"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heis_pipe -s 2 -c 1 -o gpurun_out/r02p/heis_pipe_v11 -f \
    python profiles/prof_run.py heis3d_512 3 > gpurun_out/r02p/ncu.log 2>&1
tail -2 gpurun_out/r02p/ncu.log
