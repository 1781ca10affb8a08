for v in 2 3 4; do
  VEGAS_NVCC_EXTRA="-DMSC_MINB=$v" python -m vegas_rs_b200.build --force >/dev/null 2>&1
  echo "MSC_MINB=$v"
  for w in ising3d_1024 ising2d_8192; do python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 --no-also --workload $w 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], d['value'], d['ms_per_step'])"; done
done
