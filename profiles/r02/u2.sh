#!/bin/bash
# 2 GPUs, one process each: bit-identity of the slab decomposition per kernel (default kernels, the pipelined sc kernel,
# the wave-ordered fcc/bcc kernel), the slab-group Machine, and the bench lines at N = 2
N=${1:-2}
out=gpurun_out/r02u2; mkdir -p $out
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515"
for k in ising heisenberg fcc bcc; do $T tests/mp_slab_check.py $k 2>&1 | grep -E "mp_slab_check|Error|error" | head -3; done
VEGAS_TUNE=heis_pipe=1 $T tests/mp_slab_check.py heisenberg 2>&1 | grep -E "mp_slab_check|Error|error" | head -3
VEGAS_TUNE=basis_wave=1,basis_wave_lag=1 $T tests/mp_slab_check.py fcc 2>&1 | grep -E "mp_slab_check|Error|error" | head -3
VEGAS_TUNE=basis_wave=1,basis_wave_lag=1,basis_wave_grid=40 $T tests/mp_slab_check.py bcc 2>&1 | grep -E "mp_slab_check|Error|error" | head -3
for k in ising heisenberg; do $T tests/mp_machine_check.py $k 2>&1 | grep -E "mp_machine_check|Error|error" | head -3; done
$T bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > $out/bench_n$N.json 2> $out/bench_n$N.err; tail -2 $out/bench_n$N.err
python - $N <<'PY'
import json,sys
f="gpurun_out/r02u2/bench_n%s.json"%sys.argv[1]
try:
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, d["n_gpus"], d["kernel_family"], "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], "e2e_machine", d["e2e_machine"] and "%.4g"%d["e2e_machine"]["value"])
    for k,v in d["also"].items():
        if "error" in v: print("  ",k,"ERROR",v["error"][:200]); continue
        print("  ",k, v.get("family"), "%.4g"%v["value"], v.get("ms_per_step") and "%.4f ms"%v["ms_per_step"], v.get("roofline") and "frac %.3f"%v["roofline"]["frac"], "e2e_machine", v.get("e2e_machine") and "%.4g"%v["e2e_machine"]["value"])
except Exception as e: print(f,"ERR",e)
PY
