"""Summarise `ncu --page source --csv` output: per SASS instruction samples / executed counts,
grouped into regions by the biggest sample counts.

    ncu -i X.ncu-rep --page source --csv | python tools/ncu_source_summary.py [top_n]
"""
import csv
import sys


def main():
    top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rows = list(csv.reader(sys.stdin))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    data = rows[hdr_i + 1:]
    insts = []
    for r in data:
        if r and r[0] in ("Kernel Name", "Address"):
            break                       # next launch in the report: keep the first only
        if len(r) < len(hdr):
            continue
        insts.append((r[col["Source"]].strip(), int(r[col["# Samples"]] or 0),
                      int(r[col["Instructions Executed"]] or 0)))
    tot_s = sum(s for _, s, _ in insts) or 1
    tot_i = sum(e for _, _, e in insts) or 1
    print("instructions: %d  samples: %d  warp-insts executed: %d" % (len(insts), tot_s, tot_i))
    print("--- top by samples (idx, samples%, executed, sass)")
    order = sorted(range(len(insts)), key=lambda i: -insts[i][1])[:top]
    for i in sorted(order):
        src, s, e = insts[i]
        print("%5d %6.2f%% %9d  %s" % (i, 100.0 * s / tot_s, e, src))
    # opcode histogram weighted by executed count
    hist = {}
    for src, s, e in insts:
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        op = op.split(".")[0]
        h = hist.setdefault(op, [0, 0])
        h[0] += e
        h[1] += s
    print("--- opcode histogram (executed%, samples%)")
    for op, (e, s) in sorted(hist.items(), key=lambda kv: -kv[1][0])[:25]:
        print("%-10s %6.2f%% %6.2f%%" % (op, 100.0 * e / tot_i, 100.0 * s / tot_s))


if __name__ == "__main__":
    main()
