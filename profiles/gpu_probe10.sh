#!/bin/bash
for v in new hminb4 hminb5 hminb7 hminb8 head; do
  cp build_variants/$v.so vegas_rs_b200/libvegas_gpu.so
  lag=5; [ $v = head ] && lag=4
  echo "== $v"; python profiles/wave_k_probe.py 1,4,$lag 1,4,$lag 2>&1 | tail -2
done
cp build_variants/new.so vegas_rs_b200/libvegas_gpu.so
