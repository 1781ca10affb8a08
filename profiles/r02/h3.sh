#!/bin/bash
out=${1:-r02h3}
mkdir -p gpurun_out/$out
bash profiles/r02/sweep.sh $out heis3d_512 30 "heis_pipe_pub=4,heis_pipe_lead=48" "heis_pipe_pub=4,heis_pipe_lead=56" "heis_pipe_pub=4,heis_pipe_lead=64" "heis_pipe_pub=4,heis_pipe_lead=80" "heis_pipe_pub=4,heis_pipe_lead=96" "heis_pipe_pub=6,heis_pipe_lead=64" "heis_pipe_pub=8,heis_pipe_lead=64" "heis_pipe_pub=8,heis_pipe_lead=96" "heis_pipe_pub=16,heis_pipe_lead=96"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/$out/bench.json 2> gpurun_out/$out/bench.err; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02h3/bench.json") if l.startswith("{")][-1])
print(json.dumps(d["also"]["ising_sc10_cfg0"].get("program")))
PY
tail -3 gpurun_out/$out/bench.err
