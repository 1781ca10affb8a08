#!/bin/bash
# ncu --set full captures of the Ising colour pass and of the opt-in fcc pipeline (second call: gpurun returns <= 64 MiB)
tag=${1:-r02z}
out=gpurun_out/${tag}b; mkdir -p $out
N="timeout 600 ncu --set full --clock-control none"
$N --import-source on -k regex:ising_msc -s 4 -c 1 -o $out/ising_msc -f python profiles/prof_run.py ising2d_8192 3 > $out/ncu_ising_msc2d.log 2>&1
mv $out/ising_msc.ncu-rep $out/ising_msc2d.ncu-rep
$N -k regex:ising_msc -s 4 -c 1 -o $out/ising_msc -f python profiles/prof_run.py ising3d_1024 3 > $out/ncu_ising_msc.log 2>&1
VEGAS_TUNE=basis_pipe=1 $N --import-source on -k regex:basis_pipe -s 1 -c 1 -o $out/basis_pipe -f python profiles/prof_run.py heis_fcc_384 2 > $out/ncu_basis_pipe.log 2>&1
ls -la $out
