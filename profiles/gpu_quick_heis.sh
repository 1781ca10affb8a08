timeout 600 python -m pytest tests -m gpu -x -q -k "heis or Heis or fused or wave or basis" 2>&1 | tail -2
for w in heis3d_512 heis_fcc_384; do
timeout 200 python bench.py --workload $w --no-also --no-cpu --e2e-steps 0 --steps 30 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$w', d['kernel_family'], '%.4g attempts/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['roofline']['kernel'])
    else: print(l.rstrip()[:300])"
done
