"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares (profiles/)."""
import collections
import csv
import sys

src, dst, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    v, u = float(r[14]), r[13]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u.startswith("u") else (v if u.startswith("m") else v * 1e3))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
with open(dst, "w") as f:
    f.write(f"# {note}\n# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes\n")
    f.write("kernel,launches,total_ms,share\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'"{k}",{a[0]},{a[1]:.3f},{a[1] / tot:.4f}\n')
print(open(dst).read())
