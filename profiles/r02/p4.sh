#!/bin/bash
# fcc pair launches: resident CTAs per SM x rows per colour switch x rows per tile
out=gpurun_out/r02p4; mkdir -p $out
cp vegas_rs_b200/libvegas_gpu.so /tmp/lib_keep.so
for n in 2 3 4; do
  cp profiles/r02/variants/libvegas_gpu_pair$n.so vegas_rs_b200/libvegas_gpu.so; touch vegas_rs_b200/libvegas_gpu.so
  echo "== BASIS_PAIR_MINB=$n"
  bash profiles/r02/sweep.sh r02p4/minb$n heis_fcc_384 15 "basis_pair=1,basis_pair_chunk=4" "basis_pair=1,basis_pair_chunk=8" "basis_pair=1,basis_pair_chunk=16" "basis_pair=1,basis_pair_rows=32,basis_pair_chunk=4" "basis_pair=1,basis_pair_rows=32,basis_pair_chunk=8" "basis_pair=1,basis_pair_rows=8,basis_pair_chunk=4" "basis_pair=1,basis_pair_rows=64,basis_pair_chunk=4"
done
cp /tmp/lib_keep.so vegas_rs_b200/libvegas_gpu.so
