#!/usr/bin/env python
"""ncu launch list of `bench.py` itself (gpu__time_duration.sum per launch) -> share of every class among the kernels of
the direct builds, to be compared with `roofline.share_of_step` of the bench line (per-launch times under ncu are
cold-cache and serialised: the SHARE must agree, not the absolute).
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv python bench.py --steps 2 --warmup 1
    python tools/ncu_bench_shares.py L.csv > profiles/r02_ncu_bench_launch_shares.txt"""
import collections
import csv
import re
import sys

NAMES = "spd"
UNIT = {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3}
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
col = {n: i for i, n in enumerate(rows[hi])}
agg = collections.OrderedDict()
other = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(rows[hi]) or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[col["Kernel Name"]]
    ms = float(r[col["Metric Value"]].replace(",", "")) * UNIT.get(r[col["Metric Unit"]], 1e-6)
    m = re.search(r"eri_class_kernel<(?:\(int\))?(\d), (?:\(int\))?(\d), (?:\(int\))?(\d), (?:\(int\))?(\d), (?:\(int\))?(\d)", name)
    if m:
        la, lb, lc, ld, epi = (int(x) for x in m.groups())
        if epi == 0:
            key = "dense fill / Schwarz table (epilogue 0)"
            d = other.setdefault(key, [0, 0.0]); d[0] += 1; d[1] += ms
            continue
        la, lb, lc, ld = (0 if x == 3 else x for x in (la, lb, lc, ld))
        la, lb = max(la, lb), min(la, lb)
        lc, ld = max(lc, ld), min(lc, ld)
        if la * (la + 1) // 2 + lb < lc * (lc + 1) // 2 + ld:
            la, lb, lc, ld = lc, ld, la, lb
        key = "(%s%s|%s%s)" % (NAMES[la], NAMES[lb], NAMES[lc], NAMES[ld])
        d = agg.setdefault(key, [0, 0.0]); d[0] += 1; d[1] += ms
    elif "screen_kernel" in name:
        d = agg.setdefault("screen_kernel", [0, 0.0]); d[0] += 1; d[1] += ms
    else:
        short = re.sub(r"\(.*", "", name)[:60]
        d = other.setdefault(short, [0, 0.0]); d[0] += 1; d[1] += ms
tot = sum(d[1] for d in agg.values())
print("# kernels of the direct Fock builds of one `bench.py --steps 2 --warmup 1` run under ncu (all builds of the run: warm-up, statistics, timed, e2e)")
print("# class (S2 variants and narrow-bra orientations folded into the reference's class), launches, total ms, share of %.1f ms" % tot)
for k, d in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-14s %6d %10.3f ms %6.2f %%" % (k, d[0], d[1], 100 * d[1] / tot))
print("# other kernels of the run")
for k, d in sorted(other.items(), key=lambda kv: -kv[1][1])[:12]:
    print("%-60s %6d %10.3f ms" % (k, d[0], d[1]))
