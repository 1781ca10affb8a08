#!/bin/bash
timeout 250 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 100 -k "basis or slab or fcc" 2>&1 | tail -6
for v in vminb1 vminb3 vminb4 vminb5 vminb6 vminb8; do
  cp build_variants/$v.so vegas_rs_b200/libvegas_gpu.so
  python profiles/fcc_probe.py $v 2>&1 | tail -1
done
cp build_variants/vminb1.so vegas_rs_b200/libvegas_gpu.so
