#!/bin/bash
# heis_pipe v12: sleep between failed mbarrier polls (consumer warps / helper warps)
out=${1:-r02h6}
mkdir -p gpurun_out/$out
timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q -k "pipe_kernel" > gpurun_out/$out/pytest.log 2>&1; tail -2 gpurun_out/$out/pytest.log
bash profiles/r02/sweep.sh $out heis3d_512 30 "heis_pipe=-1" "heis_pipe_backoff=32" "heis_pipe_backoff=64" "heis_pipe_backoff=128" "heis_pipe_backoff=256" "heis_pipe_backoff_helper=64" "heis_pipe_backoff_helper=128" "heis_pipe_backoff_helper=256" "heis_pipe_backoff=64,heis_pipe_backoff_helper=128" "heis_pipe_backoff=128,heis_pipe_backoff_helper=256" "heis_pipe=-1"
