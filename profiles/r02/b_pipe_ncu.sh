#!/bin/bash
# ncu --set full of one heis_pipe launch on 512^3 fp32 (recorded step), plus the launch list
mkdir -p gpurun_out/r02b
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heis_pipe -s 2 -c 1 -o gpurun_out/r02b/heis_pipe -f \
    python profiles/prof_run.py heis3d_512 3 > gpurun_out/r02b/ncu.log 2>&1
tail -3 gpurun_out/r02b/ncu.log
ls -la gpurun_out/r02b
