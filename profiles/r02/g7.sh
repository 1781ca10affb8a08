#!/bin/bash
# after the switch to Philox4x32-7: re-record the GPU golden trajectories, then the whole GPU suite
mkdir -p gpurun_out/r02g7
python tests/golden/make_gpu_replay_fixtures.py gpurun_out/r02g7/golden > gpurun_out/r02g7/golden.log 2>&1; tail -3 gpurun_out/r02g7/golden.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --durations=5 > gpurun_out/r02g7/pytest_gpu.log 2>&1; tail -12 gpurun_out/r02g7/pytest_gpu.log
