#!/bin/bash
# compute-sanitizer over the pair launches (single handle and in-process slabs) and the host-packed transfer
out=gpurun_out/r02san2; mkdir -p $out
S="timeout 900 compute-sanitizer --error-exitcode 7"
$S --tool memcheck python -m pytest tests/test_gpu_pipe.py tests/test_gpu_dropin.py -q -x -k "basis_pair or host_packed" > $out/memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 $out/memcheck.log
$S --tool racecheck python -m pytest tests/test_gpu_pipe.py -q -x -k "basis_pair_identical and size0 and random" > $out/racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 $out/racecheck.log
$S --tool synccheck python -m pytest tests/test_gpu_pipe.py -q -x -k "basis_pair_identical and size0 and random" > $out/synccheck.log 2>&1; echo "synccheck exit $?"; tail -3 $out/synccheck.log
