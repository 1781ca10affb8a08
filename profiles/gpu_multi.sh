#!/bin/bash
# multi-GPU visit (gpurun --gpus N): slab parity across processes, then the bench at N ranks.  bash profiles/gpu_multi.sh <tag> <N>
tag=${1:-multi}; N=${2:-2}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29501 tests/mp_slab_check.py ising 2>&1 | grep -E "mp_slab_check|Error|error" | head -5
timeout 300 $TR --master-port 29502 tests/mp_slab_check.py heisenberg 2>&1 | grep -E "mp_slab_check|Error|error" | head -5
timeout 300 $TR --master-port 29504 tests/mp_slab_check.py fcc 2>&1 | grep -E "mp_slab_check|Error|error" | head -5
for w in ising3d_1024 heis3d_512 heis_fcc_384; do
  timeout 600 $TR --master-port 29503 bench.py --gpus $N --steps 50 --warmup 5 --workload $w --no-also --e2e-steps 1 > $out/bench_${w}_n$N.json 2> $out/bench_${w}_n$N.err
  python -c "
import json,sys
d=json.loads(open('$out/bench_${w}_n$N.json').read().strip().splitlines()[-1]); print('$w N=$N', '%.4g' % d['value'], 'attempts/s', '%.4f ms/step' % d['ms_per_step'], d['config']['decomposition'], d['clocks'])" || tail -5 $out/bench_${w}_n$N.err
done
timeout 300 python bench.py --steps 50 --warmup 5 --workload ising3d_1024 --no-also --no-cpu --e2e-steps 0 > $out/bench_ising_n1.json 2>/dev/null
python -c "
import json; d=json.loads(open('$out/bench_ising_n1.json').read().strip().splitlines()[-1]); print('ising3d_1024 N=1', '%.4g' % d['value'], d['clocks'])"
