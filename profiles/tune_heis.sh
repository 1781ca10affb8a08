for v in 5 7 8; do
  VEGAS_NVCC_EXTRA="-DHEIS_MINB=$v" python -m vegas_rs_b200.build --force >/dev/null 2>&1
  echo "HEIS_MINB=$v"
  for t in heis_wave=-1 heis_wave=0; do
  VEGAS_TUNE=$t python bench.py --steps 30 --warmup 3 --no-cpu --e2e-steps 0 --no-also --workload heis3d_512 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  $t', '%.4g' % d['value'], d['ms_per_step'], d['roofline']['frac'])"
  done
done
