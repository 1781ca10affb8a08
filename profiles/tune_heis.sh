for v in 1 4 6 8; do
  VEGAS_NVCC_EXTRA="-DHEIS_MINB=$v" python -m vegas_rs_b200.build --force >/dev/null 2>&1
  echo "HEIS_MINB=$v"
  python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 --no-also --workload heis3d_512 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
timeout 300 python -m pytest tests -m gpu -x -q -k "heis or slab" 2>&1 | tail -3
