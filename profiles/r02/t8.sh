#!/bin/bash
# 8 GPUs, final tree: slab bit-identity (default kernels + heis_pipe), slab-group Machine, the bench line
N=${1:-8}
out=gpurun_out/r02t8; mkdir -p $out
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
nproc > $out/nproc.txt
for k in ising; do $T tests/mp_slab_check.py $k 2>&1 | grep -E "mp_slab_check|Error|error" | head -2 | tee -a $out/checks.txt; done
VEGAS_TUNE=basis_pair=1 $T tests/mp_slab_check.py fcc 2>&1 | grep -E "mp_slab_check|Error|error" | head -2 | tee -a $out/checks.txt
VEGAS_TUNE=heis_pipe=1 $T tests/mp_slab_check.py heisenberg 2>&1 | grep -E "mp_slab_check|Error|error" | head -2 | tee -a $out/checks.txt
$T tests/mp_machine_check.py ising 2>&1 | grep -E "mp_machine_check|Error|error" | head -2 | tee -a $out/checks.txt
$T bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > $out/bench_n$N.json 2> $out/bench_n$N.err; tail -2 $out/bench_n$N.err
python - $N <<'PY'
import json,sys
f="gpurun_out/r02t8/bench_n%s.json"%sys.argv[1]
try:
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, d["n_gpus"], d["kernel_family"], "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "e2e_machine", d["e2e_machine"] and "%.4g"%d["e2e_machine"]["value"])
    for k,v in d["also"].items():
        if "error" in v: print("  ",k,"ERROR",v["error"][:200]); continue
        print("  ",k, v.get("family"), "%.4g"%v["value"], v.get("ms_per_step") and "%.4f ms"%v["ms_per_step"], v.get("roofline") and "frac %.3f"%v["roofline"]["frac"], "e2e_machine", v.get("e2e_machine") and "%.4g"%v["e2e_machine"]["value"])
except Exception as e: print(f,"ERR",e)
PY
cat $out/nproc.txt
