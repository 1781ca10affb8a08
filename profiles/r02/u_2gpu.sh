#!/bin/bash
mkdir -p gpurun_out/r02u
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514"
VEGAS_TUNE=heis_pipe=1 $T tests/mp_slab_check.py heisenberg 2>&1 | grep -E "mp_slab_check|Error" | head -3
B="bench.py --gpus 2 --steps 30 --warmup 3 --no-also --no-cpu --e2e-steps 0 --workload heis3d_512"
for t in "heis_pipe=-1" "heis_pipe_lead=48" "heis_pipe_pub=2,heis_pipe_lead=32" "heis_pipe_pub=8,heis_pipe_lead=64"; do
VEGAS_TUNE=$t $T $B 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$t', d['kernel_family'], d['n_gpus'], '%.4g'%d['value'], '%.4f ms'%d['ms_per_step'])"
done
python bench.py --steps 30 --warmup 3 --no-also --no-cpu --e2e-steps 0 --workload heis3d_512 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=1', d['kernel_family'], '%.4g'%d['value'], '%.4f ms'%d['ms_per_step'])"
