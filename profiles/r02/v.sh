#!/bin/bash
mkdir -p gpurun_out/r02v
timeout 900 python -m pytest tests/test_gpu_pipe.py -x -q -k basis > gpurun_out/r02v/pytest.log 2>&1; tail -4 gpurun_out/r02v/pytest.log
bash profiles/r02/sweep.sh r02v heis_fcc_384 10 "basis_pipe=0" "basis_pipe=1" "basis_pipe=1,basis_pipe_lead=12" "basis_pipe=1,basis_pipe_lead=24" "basis_pipe=1,basis_pipe_pub=2,basis_pipe_lead=24"
VEGAS_TUNE=basis_pipe=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:basis_pipe -s 1 -c 1 -o gpurun_out/r02v/basis_pipe_v3 -f \
    python profiles/prof_run.py heis_fcc_384 2 > gpurun_out/r02v/ncu.log 2>&1
tail -2 gpurun_out/r02v/ncu.log
