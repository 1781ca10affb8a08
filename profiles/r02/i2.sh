#!/bin/bash
# Ising colour pass variants: parity tests of the family, then the two Ising workloads
out=${1:-r02i2}
mkdir -p gpurun_out/$out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -k "ising or msc" --timeout 300 > gpurun_out/$out/pytest.log 2>&1; tail -4 gpurun_out/$out/pytest.log
bash profiles/r02/sweep.sh $out ising3d_1024 20 "msc_full=0" "msc_full=1"
bash profiles/r02/sweep.sh ${out}_2d ising2d_8192 200 "msc_full=0" "msc_full=1"
