#!/bin/bash
# 2 GPUs: slab-group Machine check, slab checks, and the default bench line at N = 2 (all workloads in `also`)
mkdir -p gpurun_out/r02s
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
for k in ising heisenberg; do $T tests/mp_machine_check.py $k 2>&1 | grep -E "mp_machine_check|group :|single:|Error" | head -8; done
VEGAS_TUNE=heis_pipe=1 $T tests/mp_slab_check.py heisenberg 2>&1 | grep -E "mp_slab_check|Error" | head -3
$T bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02s/bench_n2.json 2> gpurun_out/r02s/bench_n2.err; tail -2 gpurun_out/r02s/bench_n2.err
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02s/bench_n1.json 2> gpurun_out/r02s/bench_n1.err; tail -2 gpurun_out/r02s/bench_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/r02s/bench_n1.json","gpurun_out/r02s/bench_n2.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["n_gpus"], d["kernel_family"], "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], "e2e_machine", d["e2e_machine"] and "%.4g"%d["e2e_machine"]["value"])
        for k,v in d["also"].items():
            if "error" in v: print("  ",k,"ERROR",v["error"][:200]); continue
            print("  ",k, v.get("family"), "%.4g"%v["value"], v.get("ms_per_step") and "%.4f ms"%v["ms_per_step"], v.get("roofline") and "frac %.3f"%v["roofline"]["frac"], "e2e_machine", v.get("e2e_machine") and "%.4g"%v["e2e_machine"]["value"])
    except Exception as e: print(f,"ERR",e)
PY
