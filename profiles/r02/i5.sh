#!/bin/bash
# Ising colour pass: how many of the 8 threshold bit-planes are built with LOP3 instead of IMAD (-DMSC_LOP_PLANES=n, one
# library per value under profiles/r02/variants/)
out=${1:-r02i5}
mkdir -p gpurun_out/$out
cp vegas_rs_b200/libvegas_gpu.so /tmp/lib_default.so
for n in 0 2 4 8; do
  cp profiles/r02/variants/libvegas_gpu_lop$n.so vegas_rs_b200/libvegas_gpu.so; touch vegas_rs_b200/libvegas_gpu.so
  echo "MSC_LOP_PLANES=$n"
  bash profiles/r02/sweep.sh ${out}/lop$n ising3d_1024 20 "msc_full=1"
  bash profiles/r02/sweep.sh ${out}/lop${n}_2d ising2d_8192 200 "msc_full=1"
done
cp /tmp/lib_default.so vegas_rs_b200/libvegas_gpu.so
