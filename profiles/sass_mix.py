"""Instruction mix of a kernel from `ncu --page source --csv` (reads the csv on stdin or argv[1]).
Prints warp-level executed instructions per opcode and the top stall lines."""
import csv, sys, collections
f = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
per_attempt = float(sys.argv[2]) if len(sys.argv) > 2 else None  # number of attempts (or words) to normalise by
rows = list(csv.reader(f))
kern = 0
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1][:110]; hdr = rows[i + 1]; i += 2
        ci = {k: hdr.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples")}
        ops = collections.Counter(); samp = []
        tot = 0; tthr = 0
        while i < len(rows) and not (rows[i] and rows[i][0] == "Kernel Name"):
            r = rows[i]; i += 1
            if len(r) <= ci["# Samples"]: continue
            src = r[ci["Source"]].strip(); n = int(r[ci["Instructions Executed"]] or 0); t = int(r[ci["Thread Instructions Executed"]] or 0)
            toks = src.split()
            op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
            op = op.split(".")[0] if not op.startswith(("LDG", "STG", "IMAD", "MUFU")) else ".".join(op.split(".")[:2])
            ops[op] += n; tot += n; tthr += t
            samp.append((int(r[ci["# Samples"]] or 0), src, n))
        print("==", name); print("warp instr", tot, "thread instr", tthr, ("per unit %.1f" % (tthr / per_attempt)) if per_attempt else "")
        for op, n in ops.most_common(28): print("  %-14s %10d %5.1f%%" % (op, n, 100.0 * n / tot))
        samp.sort(reverse=True)
        print("  top stall lines:")
        for s, src, n in samp[:14]: print("   %6d  %s" % (s, src[:90]))
    else:
        i += 1
