"""Per-instruction stall samples of one kernel from `ncu -i X.ncu-rep --page source --csv` (read on the CPU box).

    ncu -i gpurun_out/prof.ncu-rep --page source --csv > /tmp/src.csv
    python tools/ncu_source_summary.py /tmp/src.csv [--top N] [--range LO HI]

Prints the stall-reason totals, the N hottest instructions, and (with --range, instruction indices) every
instruction of a region with its samples and dominant stall reason -- enough to see where a loop waits.
"""
import argparse
import csv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--range", type=int, nargs=2)
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    col = {n: i for i, n in enumerate(hdr)}
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    tot = {s: 0 for s in stalls}
    samples = []
    for k, r in enumerate(body):
        n = int(r[col["# Samples"]] or 0)
        per = {s: int(r[col[s]] or 0) for s in stalls}
        for s in stalls:
            tot[s] += per[s]
        samples.append((n, k, r[col["Source"]].strip(), per, int(r[col["Instructions Executed"]] or 0)))
    total = sum(n for n, *_ in samples)
    print("kernel:", rows[0][1] if rows[0] else "?")
    print("instructions: %d, samples: %d" % (len(body), total))
    print("stall totals:", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(total, 1))
                                     for s, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
    print("-- hottest instructions")
    for n, k, src, per, ex in sorted(samples, reverse=True)[:args.top]:
        top = max(per, key=per.get)
        print("%5d %5.1f%%  #%4d x%-6d %-70s %s" % (n, 100.0 * n / max(total, 1), k, ex, src[:70], top[6:]))
    if args.range:
        lo, hi = args.range
        print("-- instructions %d..%d" % (lo, hi))
        for n, k, src, per, ex in samples[lo:hi + 1]:
            top = max(per, key=per.get) if n else ""
            print("%5d  #%4d x%-6d %-70s %s" % (n, k, ex, src[:70], top[6:] if top else ""))


if __name__ == "__main__":
    main()
