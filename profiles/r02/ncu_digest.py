"""Digest of an ncu report: key metrics + stall totals + top stalled SASS lines.  usage: ncu_digest.py report.ncu-rep [ntop]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'sm__warps_active.avg.pct', 'lts__t_sector_hit_rate', 'gpu__dram_throughput.avg.pct',
        'lts__t_bytes.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fmaheavy', 'sm__inst_executed_pipe_xu.sum',
        'sm__pipe_alu_cycles_active.avg.pct', 'sm__pipe_fma_cycles_active.avg.pct', 'sm__pipe_fmaheavy_cycles_active.avg.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w) for w in want) and 'per_second' not in h and '.max' not in h and '.min' not in h:
        print(f"{h} [{u}] {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
print("SASS: warp inst", sum(f(r, "Instructions Executed") for r in data), "samples", sum(f(r, "# Samples") for r in data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
tot = {c: sum(f(r, c) for r in data) for c in stall_cols}
print("stall totals:", [(c, int(v)) for v, c in sorted(((v, c) for c, v in tot.items()), reverse=True)[:10]])
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:ntop]:
    st = sorted(((f(r, c), c) for c in stall_cols), reverse=True)[:2]
    print(r[idx["Address"]][-5:], "%-58s" % r[idx["Source"]][:58], int(f(r, "# Samples")), int(f(r, "Instructions Executed")), [(c, int(v)) for v, c in st])
def op(r):
    t = r[idx["Source"]].split()
    return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
cnt = collections.Counter()
for r in data: cnt[op(r)] += f(r, "Instructions Executed")
print("executed by opcode (M warp-inst):", [(k, round(v / 1e6, 1)) for k, v in cnt.most_common(24)])
