"""Kernel shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`).

    python tools/ncu_launch_summary.py X.csv [--first N] [--count M]

Groups launches [first, first+count) by kernel name + block/grid size: launches, total and mean device time,
share.  Per-launch times under ncu are cold-cache and serialised -- compare SHARES, not absolutes.
"""
import argparse
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=10 ** 9)
    args = ap.parse_args()
    rows = [r for r in csv.reader(l for l in open(args.csv) if l.startswith('"'))]
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    body = [r for r in rows[1:] if len(r) == len(hdr) and r[col["Metric Name"]] == "gpu__time_duration.sum"]
    body = body[args.first:args.first + args.count]
    agg = {}
    for r in body:
        name = re.sub(r"\(.*$", "", r[col["Kernel Name"]].replace("void ", ""))
        name = re.sub(r"\(int\)", "", name)
        key = (name[:90], r[col["Block Size"]], r[col["Grid Size"]])
        t = float(r[col["Metric Value"]]) / (1e3 if r[col["Metric Unit"]] == "ns" else 1.0)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    print("launches %d..%d: %d launches, %.1f us of device time" % (args.first, args.first + len(body), len(body), total))
    for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%5.1f%% %8.1f us %5d x %7.2f us  %s block=%s grid=%s" % (100 * t / total, t, n, t / n, key[0], key[1], key[2]))


if __name__ == "__main__":
    main()
