#!/bin/bash
# host-packed State transfer: parity tests, then the e2e leg of the bench (bitmap path against the byte-per-spin path)
out=gpurun_out/r02e1; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -q -x -k "host_packed or step_host or upload or round_trip or full_size_ising" > $out/pytest.log 2>&1; tail -4 $out/pytest.log
for t in "host_pack_min=-1" "host_pack_min=0" "host_pack_chunk=16777216" "host_pack_chunk=268435456"; do
  VEGAS_TUNE=$t timeout 600 python bench.py --steps 10 --warmup 3 --no-also --no-cpu --e2e-steps 5 > "$out/bench_$t.json" 2> "$out/bench_$t.err"
  python - "$out/bench_$t.json" "$t" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d["e2e"].get("transfer","")[:40])
except Exception as e: print(sys.argv[2], "ERR", e)
PY
done
for th in 4 8 32; do
  VEGAS_HOST_THREADS=$th timeout 600 python bench.py --steps 10 --warmup 3 --no-also --no-cpu --e2e-steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threads $th e2e %.4g'%d['e2e']['value'])"
done
