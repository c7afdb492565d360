"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total us, share."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row.get("Metric Unit", "")
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else (v * 1e6 if u == "s" else v))
    n = row["Kernel Name"][:int(sys.argv[2]) if len(sys.argv) > 2 else 64]
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print("total us %.0f" % tot)
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print("  %-64s %5d %9.0f %5.1f%%" % (n, c, t, 100 * t / tot))
