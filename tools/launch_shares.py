"""Per-kernel share of a step from an ncu launch list (gpu__time_duration.sum CSV).
   python tools/launch_shares.py launches.csv[.gz] [first_launch_fraction]"""
import collections, csv, gzip, sys
path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
lines = [l for l in op(path, "rt") if not l.startswith("==")]
rows = [(x["Kernel Name"].split("(")[0].replace("wcmc::", ""), float(x["Metric Value"].replace(",", "")) / 1e3)
        for x in csv.DictReader(lines) if x.get("Metric Name") == "gpu__time_duration.sum"]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in rows:
    agg[k[:72]][0] += 1
    agg[k[:72]][1] += v
tot = sum(v[1] for v in agg.values())
print("launches %d, total %.1f us (cold-cache, serialised: compare shares, not absolutes)" % (len(rows), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-72s %5d %10.1f us %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
