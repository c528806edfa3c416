"""Summarise an ncu launch-list CSV (--metrics gpu__time_duration.sum) into a markdown table.
Usage: python tools/launch_list.py launches.csv [first_launch] [count] > profiles/xxx.md"""
import csv, re, sys
path = sys.argv[1]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    rows.append((re.sub(r"\(.*", "", r[ki]).replace("pet::", ""), ms))
rows = rows[first:first + count]
agg = {}
for k, ms in rows:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% |" % (k, n, ms, 100 * ms / tot))
print("\nTotal %.1f ms over %d launches." % (tot, len(rows)))
