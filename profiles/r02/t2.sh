#!/bin/bash
N=2
out=gpurun_out/r02t2; mkdir -p $out
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523"
$T bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > $out/bench_n$N.json 2> $out/bench_n$N.err; tail -2 $out/bench_n$N.err
$T tests/mp_machine_check.py heisenberg 2>&1 | grep -E "mp_machine_check|Error|error" | head -2
python - <<'PY'
import json
f="gpurun_out/r02t2/bench_n2.json"
try:
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, d["n_gpus"], d["kernel_family"], "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "e2e_machine", d["e2e_machine"] and "%.4g"%d["e2e_machine"]["value"])
    for k,v in d["also"].items():
        if "error" in v: print("  ",k,"ERROR",v["error"][:200]); continue
        print("  ",k, v.get("family"), "%.4g"%v["value"], v.get("ms_per_step") and "%.4f ms"%v["ms_per_step"], v.get("roofline") and "frac %.3f"%v["roofline"]["frac"], "e2e_machine", v.get("e2e_machine") and "%.4g"%v["e2e_machine"]["value"])
except Exception as e: print(f,"ERR",e)
PY
