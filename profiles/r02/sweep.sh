#!/bin/bash
# usage: sweep.sh <outdir> <workload> <steps> tune1 tune2 ...   (bench lines for several VEGAS_TUNE settings)
out=gpurun_out/$1; wl=$2; steps=$3; shift 3
mkdir -p $out
B="timeout 300 python bench.py --workload $wl --steps $steps --warmup 3 --no-also --no-cpu --e2e-steps 0"
for t in "$@"; do
  VEGAS_TUNE=$t $B > "$out/bench_$t.json" 2> "$out/bench_$t.err"
  python - "$out/bench_$t.json" "$t" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], d["kernel_family"], "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print(sys.argv[2], "ERR", e)
PY
done
