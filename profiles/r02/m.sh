#!/bin/bash
mkdir -p gpurun_out/r02m
timeout 900 python -m pytest tests/test_gpu_pipe.py -x -q > gpurun_out/r02m/pytest.log 2>&1; tail -5 gpurun_out/r02m/pytest.log
bash profiles/r02/sweep.sh r02m heis_fcc_384 10 "basis_pipe=0" "basis_pipe=1" "basis_pipe=1,basis_pipe_lead=12" "basis_pipe=1,basis_pipe_lead=24" "basis_pipe=1,basis_pipe_pub=2" "basis_pipe=1,basis_pipe_lead=32,basis_pipe_pub=2"
bash profiles/r02/sweep.sh r02m heis3d_512 20 "heis_pipe=-1"
VEGAS_TUNE=basis_pipe=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:basis_pipe -s 1 -c 1 -o gpurun_out/r02m/basis_pipe_v1 -f \
    python profiles/prof_run.py heis_fcc_384 2 > gpurun_out/r02m/ncu.log 2>&1
tail -2 gpurun_out/r02m/ncu.log
