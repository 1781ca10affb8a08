#!/bin/bash
# basis_pair as the default for fcc lattices beyond L2: whole GPU suite, fcc bench, ncu capture of one pair launch
out=gpurun_out/r02p5; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
bash profiles/r02/sweep.sh r02p5 heis_fcc_384 30 "basis_pair=-1" "basis_pair=0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heis_basis_pair -s 2 -c 1 -o $out/basis_pair -f python profiles/prof_run.py heis_fcc_384 3 > $out/ncu.log 2>&1; tail -1 $out/ncu.log
