#!/bin/bash
# 2 GPUs: new parity tests on GPU 0; slab checks (multi-process, CUDA IPC) for every family incl. the pipelined kernel;
# weak-scaling bench lines for heis3d_512 (pipe slab) at N = 2
mkdir -p gpurun_out/r02l
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_statistics.py -x -q -s > gpurun_out/r02l/pytest_new.log 2>&1; tail -8 gpurun_out/r02l/pytest_new.log
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for k in ising heisenberg fcc; do $T tests/mp_slab_check.py $k 2>&1 | grep mp_slab_check; done
VEGAS_TUNE=heis_pipe=1 $T tests/mp_slab_check.py heisenberg 2>&1 | grep -E "mp_slab_check|Error|error" | head -5
B="bench.py --gpus 2 --steps 20 --warmup 3 --no-also --no-cpu --e2e-steps 0 --workload heis3d_512"
$T $B > gpurun_out/r02l/bench_heis_n2.json 2> gpurun_out/r02l/bench_heis_n2.err; tail -c 600 gpurun_out/r02l/bench_heis_n2.json; tail -3 gpurun_out/r02l/bench_heis_n2.err
VEGAS_TUNE=heis_pipe=0 $T $B > gpurun_out/r02l/bench_heis_n2_old.json 2> gpurun_out/r02l/bench_heis_n2_old.err; tail -c 300 gpurun_out/r02l/bench_heis_n2_old.json
python bench.py --steps 20 --warmup 3 --no-also --no-cpu --e2e-steps 0 --workload heis3d_512 > gpurun_out/r02l/bench_heis_n1.json 2>&1; tail -c 300 gpurun_out/r02l/bench_heis_n1.json
