#!/bin/bash
# fcc pair launches with 4 / 2 resident CTAs per SM (128 / 221 registers: more loads in flight per thread, a smaller window in L2)
out=gpurun_out/r02p3; mkdir -p $out
cp vegas_rs_b200/libvegas_gpu.so /tmp/lib_keep.so
for n in 4 2; do
  cp profiles/r02/variants/libvegas_gpu_pair$n.so vegas_rs_b200/libvegas_gpu.so; touch vegas_rs_b200/libvegas_gpu.so
  echo "== BASIS_PAIR_MINB=$n"
  bash profiles/r02/sweep.sh r02p3/minb$n heis_fcc_384 20 "basis_pair=1,basis_pair_chunk=1" "basis_pair=1,basis_pair_chunk=2" "basis_pair=1,basis_pair_chunk=4" "basis_pair=1,basis_pair_rows=32,basis_pair_chunk=2"
  VEGAS_TUNE=basis_pair=1,basis_pair_chunk=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:heis_basis_pair -s 2 -c 1 --csv python profiles/prof_run.py heis_fcc_384 3 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{printf "%s=%s%s ", $(NF-2), $NF, $(NF-1)}' | tr -d '"'; echo
done
cp /tmp/lib_keep.so vegas_rs_b200/libvegas_gpu.so
