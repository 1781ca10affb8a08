#!/bin/bash
# fcc pair launches, colours alternating in chunks of rows: parity, bench by chunk / tile rows, traffic
out=gpurun_out/r02p2; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_pipe.py -q -x -k "basis_pair" > $out/pytest.log 2>&1; tail -3 $out/pytest.log
bash profiles/r02/sweep.sh r02p2 heis_fcc_384 20 "basis_pair=0" "basis_pair=1" "basis_pair_chunk=1" "basis_pair_chunk=2" "basis_pair_chunk=8" "basis_pair_rows=32,basis_pair_chunk=4" "basis_pair_rows=64,basis_pair_chunk=4" "basis_pair_rows=32,basis_pair_chunk=2" "basis_pair_rows=8,basis_pair_chunk=4"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:heis_basis_pair -s 2 -c 2 --csv python profiles/prof_run.py heis_fcc_384 3 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{printf "%s=%s%s\n", $(NF-2), $NF, $(NF-1)}' | tr -d '"'
