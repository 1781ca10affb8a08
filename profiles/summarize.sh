#!/bin/bash
# Summarise an ncu report into profiles/<name>.metrics.txt (+ instruction mix): bash profiles/summarize.sh gpurun_out/r01a/heis_stencil.ncu-rep r01a_heis_stencil [units-per-launch]
rep=$1; name=$2; units=$3
ncu -i $rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
keys=['Kernel Name','launch__grid_size','launch__block_size','launch__registers_per_thread','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__cycles_elapsed.avg','smsp__cycles_active.avg']
for r in rows[2:]:
    print('---')
    for k in keys:
        if k in hdr: print('%-66s %s %s' % (k, r[hdr.index(k)], rows[1][hdr.index(k)]))
" > profiles/$name.metrics.txt
ncu -i $rep --page source --csv 2>/dev/null > /tmp/$name.src.csv
python profiles/sass_mix.py /tmp/$name.src.csv $units > profiles/$name.sassmix.txt
echo "wrote profiles/$name.metrics.txt profiles/$name.sassmix.txt"
