"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and
shares.  usage: python profiles/summarize_launches.py launches.csv [skip_first_n] > profiles/<name>.txt"""
import collections
import csv
import re
import sys


def main(path, skip=0):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    n = 0
    for r in rows[1:]:
        if len(r) <= ival or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        n += 1
        if n <= skip:
            continue
        v = float(r[ival].replace(",", ""))
        u = r[iunit]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
        name = re.sub(r"\(.*", "", r[iname])[:110]
        c = tot.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += us
    all_us = sum(v[1] for v in tot.values())
    print("launches counted: %d (skipped first %d); total %.3f ms" % (n - skip, skip, all_us / 1e3))
    for k, (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%7.3f ms  %5.1f %%  x%-4d %s" % (us / 1e3, 100 * us / all_us, cnt, k))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
