for v in "heis_wave=0" "heis_wave=-1" "heis_wave_planes=8,heis_wave_lag=4" "heis_wave_planes=8,heis_wave_lag=2" "heis_wave_planes=4,heis_wave_lag=8" "heis_wave_planes=2,heis_wave_lag=8" "heis_wave_planes=6,heis_wave_lag=4"; do
  VEGAS_TUNE="$v" timeout 120 python bench.py --workload heis3d_512 --no-also --no-cpu --e2e-steps 0 --steps 30 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$v', '%.4g attempts/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])
    else: print(l.rstrip()[:300])"
done
