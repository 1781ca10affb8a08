for v in 1 2 4; do
  VEGAS_NVCC_EXTRA="-DMSC_ROWS_2D=$v" python -m vegas_rs_b200.build --force >/dev/null 2>&1 || echo BUILD FAILED
  python bench.py --steps 200 --warmup 5 --no-cpu --e2e-steps 0 --no-also --workload ising2d_8192 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MSC_ROWS_2D=$v', '%.4g' % d['value'], d['ms_per_step'])"
done
python -m vegas_rs_b200.build --force >/dev/null 2>&1
timeout 300 python -m pytest tests -m gpu -x -q -k "ising or onsager or exact" 2>&1 | tail -2
