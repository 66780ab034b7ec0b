"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg, tot, seq = collections.OrderedDict(), 0.0, []
for row in csv.DictReader(lines):
    name = row["Kernel Name"]; v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    v = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
    m = re.search(r"(gemv_\w*<[^>]*>|attn_\w+<[^>]*>|sampler_kernel|\w+_kernel)", name)
    key = m.group(1) if m else name[:60]
    seq.append((key, v, row.get("Grid Size", "")))
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"total {tot:.1f} us over {len(seq)} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:46s} n={n:4d} total={t:9.1f}us avg={t/n:7.2f}us share={t/tot*100:5.1f}%")
if len(sys.argv) > 2:
    for s in seq[: int(sys.argv[2])]:
        print(s)
