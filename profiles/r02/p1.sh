#!/bin/bash
# fcc pair launches (heis_basis_pair_kernel): parity, the 384^3 bench by rows per CTA, traffic
out=gpurun_out/r02p1; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_host_layer.py -q -x -k "basis or fcc or bcc" > $out/pytest.log 2>&1; tail -5 $out/pytest.log
bash profiles/r02/sweep.sh r02p1 heis_fcc_384 20 "basis_pair=0" "basis_pair=1" "basis_pair_rows=8" "basis_pair_rows=12" "basis_pair_rows=24" "basis_pair_rows=32" "basis_pair_rows=48"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:heis_basis_pair -s 2 -c 2 --csv python profiles/prof_run.py heis_fcc_384 3 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{printf "%s=%s%s\n", $(NF-2), $NF, $(NF-1)}' | tr -d '"'
