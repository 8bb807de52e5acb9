"""Per-instruction dump of an `ncu --page source --csv` stream with stall columns.
    ncu -i X.ncu-rep --page source --csv | python tools/ncu_regions.py [min_samples]
"""
import csv
import sys

names = ['stall_barrier', 'stall_branch_resolving', 'stall_dispatch', 'stall_drain', 'stall_lg', 'stall_long_sb',
         'stall_math', 'stall_membar', 'stall_mio', 'stall_misc', 'stall_no_inst', 'stall_not_selected',
         'stall_selected', 'stall_short_sb', 'stall_sleep', 'stall_tex', 'stall_wait']
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if r and r[0] in ('Kernel Name', 'Address'):
        break
    data.append(r)
mins = int(sys.argv[1]) if len(sys.argv) > 1 else 0
tot = sum(int(r[col['# Samples']] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {n: sum(int(r[col[n]] or 0) for r in data) for n in names}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for i, r in enumerate(data):
    s = int(r[col['# Samples']] or 0)
    if s < mins:
        continue
    st = {n[6:]: int(r[col[n]] or 0) for n in names if int(r[col[n]] or 0)}
    print(i, str(s).rjust(4), r[col['Instructions Executed']].rjust(7), r[col['L1 Wavefronts Shared']].rjust(6),
          r[col['L1 Wavefronts Shared Ideal']].rjust(6), r[col['Source']].strip().ljust(52), st)
