"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel.

usage: summarize_launches.py <launches.csv> <out_prefix>
writes <out_prefix>_summary.txt and <out_prefix>_ibss_rows.csv
"""
import collections, csv, re, sys

src, out = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
keep = [hdr]
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn])
    v = float(r[mv].replace(",", ""))
    ms = v / 1e6 if r[mu].startswith("n") else (v / 1e3 if r[mu].startswith("u") else v)
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += ms
    a[2] = max(a[2], ms)
    if "ibss_kernel" in name:
        keep.append(r)
tot = sum(a[1] for a in agg.values())
with open(out + "_summary.txt", "w") as fh:
    fh.write(f"# source: {src}  (ncu --metrics gpu__time_duration.sum --clock-control none; serialised launches)\n")
    fh.write(f"# total GPU time {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches\n")
    fh.write("launches      total_ms        max_ms   share  kernel\n")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"{a[0]:8d} {a[1]:13.2f} {a[2]:13.2f} {100 * a[1] / tot:6.2f}%  {n}\n")
csv.writer(open(out + "_ibss_rows.csv", "w")).writerows(keep)
print(open(out + "_summary.txt").read())
