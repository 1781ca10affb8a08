#!/bin/bash
# basis_wave (K4w): parity tests, then the fcc 384^3 bench over lag / items per thread
out=${1:-r02w}
mkdir -p gpurun_out/$out
timeout 900 python -m pytest tests/test_gpu_pipe.py -x -q -k "basis_wave" > gpurun_out/$out/pytest.log 2>&1; tail -6 gpurun_out/$out/pytest.log
bash profiles/r02/sweep.sh $out heis_fcc_384 10 "basis_wave=0" "basis_wave=1" "basis_wave_lag=1" "basis_wave_lag=3" "basis_wave_lag=4" "basis_wave_lag=2,basis_wave_ipt=2" "basis_wave_lag=3,basis_wave_ipt=2" "basis_wave_lag=4,basis_wave_ipt=2" "basis_wave_lag=4,basis_wave_ipt=4"
VEGAS_TUNE=basis_wave=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:basis_wave -s 1 -c 1 -o gpurun_out/$out/basis_wave -f \
    python profiles/prof_run.py heis_fcc_384 2 > gpurun_out/$out/ncu.log 2>&1
tail -2 gpurun_out/$out/ncu.log
