#!/bin/bash
out=gpurun_out/r02p6; mkdir -p $out
bash profiles/r02/sweep.sh r02p6 heis_fcc_384 20 "basis_pair=1" "basis_pair_chunk=2" "basis_pair_chunk=3" "basis_pair_chunk=5" "basis_pair_chunk=6" "basis_pair_rows=24" "basis_pair_rows=48" "basis_pair_rows=96" "basis_pair_rows=128,basis_pair_chunk=4" "basis_pair_rows=48,basis_pair_chunk=3" "basis_pair_rows=384,basis_pair_chunk=4"
