"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total time / share.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [first_fraction last_fraction] > profiles/xxx.md"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
        rows.append((re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "")[:90], us, x["Grid Size"], x["Block Size"]))
    return rows


def main():
    path = sys.argv[1]
    lo = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    hi = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    rows = load(path)
    n = len(rows)
    sel = rows[int(n * lo):int(n * hi)]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, us, _, _ in sel:
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    print("source: %s, launches %d..%d of %d (cold-cache, serialised ncu timings: compare SHARES)\n" %
          (path, int(n * lo), int(n * hi), n))
    print("total %.1f us over %d launches\n" % (tot, len(sel)))
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.001:
            continue
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, v[0], v[1], 100 * v[1] / tot))


if __name__ == "__main__":
    main()
