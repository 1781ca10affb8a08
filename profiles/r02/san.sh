#!/bin/bash
# compute-sanitizer over the small parity cases of the round-2 kernels (memcheck: out-of-bounds / misaligned; racecheck:
# shared-memory hazards; synccheck: barrier misuse)
out=gpurun_out/r02san; mkdir -p $out
S="timeout 900 compute-sanitizer --error-exitcode 7"
$S --tool memcheck python -m pytest tests/test_gpu_pipe.py -q -x -k "size0 or size3 or replays or opt_in or default" > $out/memcheck_pipe.log 2>&1; echo "memcheck pipe exit $?"; tail -3 $out/memcheck_pipe.log
$S --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "msc_sweep_replay or slab_decomposition or energies_bit_exact" > $out/memcheck_parity.log 2>&1; echo "memcheck parity exit $?"; tail -3 $out/memcheck_parity.log
$S --tool racecheck python -m pytest tests/test_gpu_pipe.py -q -x -k "(pipe_kernel_identical and size0 and random) or (basis_wave_identical and size0 and random) or (basis_pipe_identical and size0 and random)" > $out/racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 $out/racecheck.log
$S --tool synccheck python -m pytest tests/test_gpu_pipe.py -q -x -k "(pipe_kernel_identical and size0 and random) or (basis_wave_identical and size0 and random) or (basis_pipe_identical and size0 and random)" > $out/synccheck.log 2>&1; echo "synccheck exit $?"; tail -3 $out/synccheck.log
grep -h "ERROR SUMMARY" $out/*.log | sort | uniq -c
