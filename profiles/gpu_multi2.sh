#!/bin/bash
# 2-GPU check of the round's last changes: fcc slabs across processes on the vector kernel, bench at N=2 (default workload + also), fcc bench at N=2
out=gpurun_out/r01m; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 90 $TR --master-port 29504 tests/mp_slab_check.py fcc 2>&1 | grep -E "mp_slab_check|Error|error" | head -5 | tee $out/mp_fcc.txt
timeout 150 $TR --master-port 29503 bench.py --gpus 2 --steps 20 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench n2 exit $?"
timeout 90 $TR --master-port 29505 bench.py --gpus 2 --steps 10 --warmup 3 --workload heis_fcc_384 --no-also --no-cpu --e2e-steps 0 > $out/bench_fcc_n2.json 2> $out/bench_fcc_n2.err; echo "fcc n2 exit $?"
python - <<'PY'
import json
for f in ("bench_n2", "bench_fcc_n2"):
    try:
        d = json.loads(open(f"gpurun_out/r01m/{f}.json").read().strip().splitlines()[-1])
        print(f, "%.4g" % d["value"], "%.4f ms/step" % d["ms_per_step"], d["config"]["decomposition"], d["clocks"], {k: "%.4g" % v["value"] for k, v in d.get("also", {}).items()})
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 $out/bench_n2.err $out/bench_fcc_n2.err
