#!/usr/bin/env python
"""ncu launch list (tools/ncu_launches.sh: gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch)
-> one line per kernel group of a direct Fock build:

    python tools/ncu_dram_table.py gpurun_out/<tag>_launches.csv profiles/r02_ncu_dram_<workload>.csv

Output rows `name,dram_read_bytes,dram_write_bytes,ms,launches`; ERI class kernels are summed per angular-momentum class
under the name bench.py uses ("(ps|ss)": far + near + slow lists), the screening kernel is its own row.  bench.py reads
`roofline.traffic` from this file (and prints null when no capture of the benchmarked workload exists)."""
import collections
import csv
import re
import sys

NAMES = "spd"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3,
        "msecond": 1.0, "ms": 1.0, "second": 1e3}


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    col = {n: i for i, n in enumerate(rows[hi])}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(rows[hi]):
            continue
        name, metric, unit = r[col["Kernel Name"]], r[col["Metric Name"]], r[col["Metric Unit"]]
        val = float(r[col["Metric Value"]].replace(",", "")) * UNIT.get(unit, 1.0)
        m = re.search(r"eri_class_kernel<(?:\(int\))?(\d), (?:\(int\))?(\d), (?:\(int\))?(\d), (?:\(int\))?(\d), (?:\(int\))?(\d)", name)
        if m:
            la, lb, lc, ld, epi = (int(x) for x in m.groups())
            if epi == 0:
                continue                      # EPI_STORE launches belong to the Schwarz table, not to the build
            # shell type code 3 = S2 pseudo-shell (two s contractions): its launches count towards the plain class of the
            # members, and kernels that run with the narrow pair as the bra are listed under the reference's class name
            la, lb, lc, ld = (0 if x == 3 else x for x in (la, lb, lc, ld))
            la, lb = max(la, lb), min(la, lb)
            lc, ld = max(lc, ld), min(lc, ld)
            if la * (la + 1) // 2 + lb < lc * (lc + 1) // 2 + ld:
                la, lb, lc, ld = lc, ld, la, lb
            key = "(%s%s|%s%s)" % (NAMES[la], NAMES[lb], NAMES[lc], NAMES[ld])
        elif "screen_kernel" in name:
            key = "screen_kernel"
        else:
            continue
        d = agg.setdefault(key, [0.0, 0.0, 0.0, 0])
        if metric == "dram__bytes_read.sum":
            d[0] += val
        elif metric == "dram__bytes_write.sum":
            d[1] += val
        elif metric == "gpu__time_duration.sum":
            d[2] += val
            d[3] += 1
    with open(dst, "w") as f:
        f.write("# name,dram_read_bytes,dram_write_bytes,ms,launches  (ncu --clock-control none, cold serialised launches, one build)\n")
        for k, d in agg.items():
            f.write("%s,%.0f,%.0f,%.4f,%d\n" % (k, d[0], d[1], d[2], d[3]))
    tot = [sum(d[i] for d in agg.values()) for i in range(3)]
    print("total dram read %.3f GB write %.3f GB, kernel time %.2f ms" % (tot[0] / 1e9, tot[1] / 1e9, tot[2]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
