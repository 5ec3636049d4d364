#!/usr/bin/env python
"""Execution-weighted opcode histogram + hottest instructions from an `ncu --page source --csv` export (gz)."""
import collections, csv, gzip, io, re, sys
rows = list(csv.reader(io.StringIO(gzip.open(sys.argv[1], "rt").read())))[2:]
tot_exec = tot_samp = 0
by_op = collections.Counter(); samp_op = collections.Counter(); levels = collections.Counter()
recs = []
for r in rows:
    try:
        src, samp, execd = r[1].strip(), int(r[2]), int(r[5])
    except Exception:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else src.split()[0]
    by_op[op] += execd; samp_op[op] += samp; tot_exec += execd; tot_samp += samp
    levels[execd] += 1
    recs.append((samp, execd, src))
print("total warp-instructions executed %.3e, samples %d, static instructions %d" % (tot_exec, tot_samp, len(recs)))
print("-- opcode: executed share / stall-sample share")
for op, n in by_op.most_common(22):
    print("  %-10s %6.2f%%  %6.2f%%" % (op, 100.0 * n / tot_exec, 100.0 * samp_op[op] / max(tot_samp, 1)))
print("-- execution-count levels (count: #static instr)")
for lv, n in sorted(levels.items(), key=lambda kv: -kv[0] * kv[1])[:10]:
    print("  executed %d x : %d instructions -> %.1f%% of dynamic" % (lv, n, 100.0 * lv * n / tot_exec))
print("-- hottest instructions by stall samples")
for samp, execd, src in sorted(recs, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("  %6d  exec %9d  %s" % (samp, execd, src[:90]))
