#!/bin/bash
# ncu --set full of basis_wave (K4w) on fcc 384^3
mkdir -p gpurun_out/r02x
VEGAS_TUNE=basis_wave_lag=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:basis_wave -s 1 -c 1 -o gpurun_out/r02x/basis_wave_v1 -f \
    python profiles/prof_run.py heis_fcc_384 2 > gpurun_out/r02x/ncu.log 2>&1
tail -2 gpurun_out/r02x/ncu.log
