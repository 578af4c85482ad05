"""profiles/ncu_traffic.json from an ncu launch list (gpu__time_duration + dram bytes per launch).
   python tools/traffic_from_launches.py profiles/rNN_ncu_launches_step.csv.gz > profiles/ncu_traffic.json"""
import collections, csv, gzip, json, sys
path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
lines = [l for l in op(path, "rt") if not l.startswith("==")]
per = collections.defaultdict(dict)
for x in csv.DictReader(lines):
    per[x["ID"]]["name"] = x["Kernel Name"]
    per[x["ID"]][x["Metric Name"]] = (float(x["Metric Value"].replace(",", "")), x["Metric Unit"])
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, n = 0.0, 0
for d in per.values():
    if "conv_igemm_kernel" not in d.get("name", ""):
        continue
    b = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v, u = d.get(m, (0.0, "byte"))
        b += v * scale.get(u, 1.0)
    tot += b
    n += 1
print(json.dumps({"conv_igemm_k5_k3_bytes_per_launch": int(tot / max(n, 1)),
                  "conv_igemm_bytes_per_launch": int(tot / max(n, 1)),
                  "source": "%s (dram__bytes_read.sum + dram__bytes_write.sum, mean over the %d conv_igemm_kernel "
                            "launches of one step; all of them are 5x5 / 3x3 layers)" % (path, n)}, indent=1))
