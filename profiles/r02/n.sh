#!/bin/bash
mkdir -p gpurun_out/r02n
timeout 900 python -m pytest tests/test_gpu_pipe.py -x -q -k basis > gpurun_out/r02n/pytest.log 2>&1; tail -3 gpurun_out/r02n/pytest.log
bash profiles/r02/sweep.sh r02n heis_fcc_384 10 "basis_pipe=1" "basis_pipe=1,basis_pipe_lead=24" "basis_pipe=1,basis_pipe_pub=2"
VEGAS_TUNE=basis_pipe=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:basis_pipe -s 1 -c 1 -o gpurun_out/r02n/basis_pipe_v2 -f \
    python profiles/prof_run.py heis_fcc_384 2 > gpurun_out/r02n/ncu.log 2>&1
tail -2 gpurun_out/r02n/ncu.log
