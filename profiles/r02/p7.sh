#!/bin/bash
# pair launches on z-slabs: in-process slab tests (1 GPU), then 2 GPUs across processes + bench
N=${1:-2}
out=gpurun_out/r02p7; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_parity.py -q -x -k "basis_pair or slab" > $out/pytest.log 2>&1; tail -4 $out/pytest.log
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
for k in fcc bcc; do $T tests/mp_slab_check.py $k 2>&1 | grep -E "mp_slab_check|Error|error" | head -3; done
VEGAS_TUNE=basis_pair=1 $T tests/mp_slab_check.py fcc 2>&1 | grep -E "mp_slab_check|Error|error" | head -3
$T bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --workload heis_fcc_384 --no-also --e2e-steps 1 > $out/bench_fcc_n$N.json 2> $out/bench_fcc_n$N.err; tail -2 $out/bench_fcc_n$N.err
python - $N <<'PY'
import json,sys
f="gpurun_out/r02p7/bench_fcc_n%s.json"%sys.argv[1]
try:
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, d["n_gpus"], d["kernel_family"], "%.4g"%d["value"], "%.4f ms"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e_machine", d["e2e_machine"] and "%.4g"%d["e2e_machine"]["value"])
except Exception as e: print(f,"ERR",e)
PY
