"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/launch_summary.py launches.csv [first_launch_id [last_launch_id]]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, vi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    try:
        v, i = float(r[vi].replace(",", "")), int(r[ii])
    except ValueError:
        continue
    if lo <= i <= hi:
        name = r[ki].split("(")[0][-60:]
        tot[name] += v
        cnt[name] += 1
total = sum(tot.values())
print(f"# launches {sum(cnt.values())}  total {total / 1e6:.3f} ms")
for k, v in tot.most_common(25):
    print(f"{v / 1e6:10.3f} ms {100 * v / total:5.1f}%  n={cnt[k]:4d}  {k}")
