#!/bin/bash
# heis_pipe after Philox-7: ring / lead / publication sweep; cfg0 program leg of the bench
out=${1:-r02h2}
mkdir -p gpurun_out/$out
bash profiles/r02/sweep.sh $out heis3d_512 30 "heis_pipe=-1" "heis_pipe_pub=2,heis_pipe_lead=24" "heis_pipe_pub=2,heis_pipe_lead=32" "heis_pipe_pub=4,heis_pipe_lead=24" "heis_pipe_pub=4,heis_pipe_lead=40" "heis_pipe_pub=4,heis_pipe_lead=48" "heis_pipe_pub=8,heis_pipe_lead=48" "heis_pipe_pub=3,heis_pipe_lead=32" "heis_pipe_stages=5,heis_pipe_own=3" "heis_pipe_stages=6,heis_pipe_own=4"
timeout 600 python bench.py --workload ising_sc10_cfg0 > gpurun_out/$out/bench_cfg0.json 2> gpurun_out/$out/bench_cfg0.err; tail -c 900 gpurun_out/$out/bench_cfg0.json; tail -3 gpurun_out/$out/bench_cfg0.err
