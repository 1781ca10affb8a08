timeout 600 python -m pytest tests -m gpu -x -q -k "basis_lattices or heisenberg_energies or sweep_replay or langevin or statistics or host or machine" 2>&1 | tail -4
timeout 300 python bench.py --workload heis_fcc_384 --no-also --no-cpu --e2e-steps 0 --steps 10 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('fcc384', d['kernel_family'], '%.4g attempts/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['gpu_launches'])
    else: print(l.rstrip()[:300])"
if [ "$1" = "ncu" ]; then
 timeout 300 ncu --set full --clock-control none --import-source on -k regex:heis_basis -s 2 -c 2 -f -o gpurun_out/fcc python profiles/prof_run.py heis_fcc_384 2 > gpurun_out/fcc_ncu.log 2>&1; tail -2 gpurun_out/fcc_ncu.log
fi
