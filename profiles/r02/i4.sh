#!/bin/bash
out=${1:-r02i4}
mkdir -p gpurun_out/$out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ising_msc -s 4 -c 1 -o gpurun_out/$out/ising_msc -f python profiles/prof_run.py ising3d_1024 3 > gpurun_out/$out/ncu.log 2>&1
tail -2 gpurun_out/$out/ncu.log
