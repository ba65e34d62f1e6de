"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum[,more] --csv --log-file X.csv ...).

    python tools/launch_summary.py X.csv [--by-metric]   -> kernel, launches, total ns, share, min..max ns per launch
With --by-metric every collected metric is averaged per kernel name (used for the C4 counter table).
"""
import csv
import sys
from collections import OrderedDict, defaultdict


def rows_of(path):
    hdr = None
    for r in csv.reader(open(path, errors="replace")):
        if hdr is None:
            if len(r) > 10 and r[0] == "ID":
                hdr = r
            continue
        if len(r) == len(hdr):
            yield dict(zip(hdr, r))


def main():
    path = sys.argv[1]
    by_metric = "--by-metric" in sys.argv
    per = OrderedDict()
    for r in rows_of(path):
        key = (r["ID"], r["Kernel Name"])
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        per.setdefault(key, {})[r["Metric Name"]] = (v, r["Metric Unit"])
    agg = OrderedDict()
    for (_, name), m in per.items():
        agg.setdefault(name, []).append(m)
    if not by_metric:
        tot = sum(m["gpu__time_duration.sum"][0] for ms in agg.values() for m in ms if "gpu__time_duration.sum" in m)
        print("# kernel, launches, total ns, share, min..max ns per launch")
        for name, ms in agg.items():
            t = [m["gpu__time_duration.sum"][0] for m in ms if "gpu__time_duration.sum" in m]
            print(f"{name[:92]:92s} {len(t):6d} {sum(t):14.0f} {100 * sum(t) / tot:6.2f}%  {min(t):.0f}..{max(t):.0f}")
        return
    for name, ms in agg.items():
        print(f"## {name[:150]}  ({len(ms)} launches; per-launch mean, and min..max of the duration)")
        keys = defaultdict(list)
        units = {}
        for m in ms:
            for k, (v, u) in m.items():
                keys[k].append(v)
                units[k] = u
        for k, vs in keys.items():
            extra = f"   [{min(vs):.4g} .. {max(vs):.4g}]" if k == "gpu__time_duration.sum" else ""
            print(f"  {k:72s} {sum(vs) / len(vs):16.6g} {units[k]}{extra}")


if __name__ == "__main__":
    main()
