#!/bin/bash
# resident kernel v2 tests + probe; fcc kernel register-cap variants (A/B on one box)
tag=${1:-r01y}
out=gpurun_out/$tag; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_resident.py tests/test_host_layer.py -m gpu -q --timeout 120 > $out/pytest_resident.log 2>&1; echo "pytest exit $?" >> $out/pytest_resident.log
timeout 60 python profiles/resident_probe.py > $out/resident_probe.txt 2>&1
for v in minb1 minb12 minb16; do
  cp build_variants/$v.so vegas_rs_b200/libvegas_gpu.so
  timeout 60 python profiles/fcc_probe.py $v >> $out/fcc_probe.txt 2>&1
done
cp build_variants/minb12.so vegas_rs_b200/libvegas_gpu.so
tail -6 $out/pytest_resident.log; cat $out/resident_probe.txt $out/fcc_probe.txt
