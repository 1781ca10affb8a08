#!/bin/bash
# fcc colour launches: own sublattice read / written with evict-first accesses (-DBASIS_STREAM=1) against the default
out=gpurun_out/r02f1; mkdir -p $out
cp vegas_rs_b200/libvegas_gpu.so /tmp/lib_default.so
for v in default stream1 default stream1; do
  if [ $v = default ]; then cp /tmp/lib_default.so vegas_rs_b200/libvegas_gpu.so; else cp profiles/r02/variants/libvegas_gpu_$v.so vegas_rs_b200/libvegas_gpu.so; fi
  touch vegas_rs_b200/libvegas_gpu.so
  echo "== $v"
  bash profiles/r02/sweep.sh r02f1/$v heis_fcc_384 20 "basis_vec=1"
done
cp profiles/r02/variants/libvegas_gpu_stream1.so vegas_rs_b200/libvegas_gpu.so; touch vegas_rs_b200/libvegas_gpu.so
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:heis_basis_vec -s 4 -c 4 --csv python profiles/prof_run.py heis_fcc_384 3 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{printf "%s=%s%s\n", $(NF-2), $NF, $(NF-1)}' | tr -d '"'
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "basis" 2>&1 | tail -2
cp /tmp/lib_default.so vegas_rs_b200/libvegas_gpu.so
