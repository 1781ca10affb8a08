#!/bin/bash
# speed probe (no parity: the oracle still draws 10 rounds): the four bench workloads with a library built with -DVEGAS_PHILOX_ROUNDS=7
out=${1:-r02y}
mkdir -p gpurun_out/$out
for wl in ising3d_1024 ising2d_8192 heis3d_512 heis_fcc_384; do
  bash profiles/r02/sweep.sh $out $wl 20 "resident_max=8192"
  mv "gpurun_out/$out/bench_resident_max=8192.json" gpurun_out/$out/bench_$wl.json
done
