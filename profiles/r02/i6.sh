#!/bin/bash
# Ising colour pass compiled for 5 CTAs per SM (48 registers, 16-40 B of spills) against the default 4 (64 registers)
out=gpurun_out/r02i6; mkdir -p $out
cp vegas_rs_b200/libvegas_gpu.so /tmp/lib_keep.so
for v in default minb5 default minb5; do
  if [ $v = default ]; then cp /tmp/lib_keep.so vegas_rs_b200/libvegas_gpu.so; else cp profiles/r02/variants/libvegas_gpu_$v.so vegas_rs_b200/libvegas_gpu.so; fi
  touch vegas_rs_b200/libvegas_gpu.so
  echo "== $v"
  bash profiles/r02/sweep.sh r02i6/$v ising3d_1024 30 "msc_full=1"
  bash profiles/r02/sweep.sh r02i6/${v}_2d ising2d_8192 300 "msc_full=1"
done
cp /tmp/lib_keep.so vegas_rs_b200/libvegas_gpu.so
