out=gpurun_out/${1:-h2p}; mkdir -p $out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:heis_stencil -s 4 -c 2 -f -o $out/heis2p python profiles/prof_run.py heis3d_512 4 > $out/ncu.log 2>&1; tail -2 $out/ncu.log
