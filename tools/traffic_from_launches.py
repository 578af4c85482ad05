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
agg = collections.defaultdict(lambda: [0, 0.0])
for d in per.values():
    name = d.get("name", "").split("(")[0].replace("void ", "").replace("wcmc::", "").strip()
    b = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v, u = d.get(m, (0.0, "byte"))
        b += v * scale.get(u, 1.0)
    agg[name][0] += 1
    agg[name][1] += b
ours = {k: {"launches": v[0], "dram_bytes_per_launch": int(v[1] / v[0])} for k, v in sorted(agg.items())
        if any(t in k for t in ("conv_", "pathnet", "kernel_apply", "fmse", "adam", "slab", "wgrad"))}
k5 = ours.get("conv_igemm_kernel<1, 5, 0>", ours.get("conv_igemm_kernel<1, 5>", {"dram_bytes_per_launch": None, "launches": 0}))
print(json.dumps({"conv_igemm_k5_bytes_per_launch": k5["dram_bytes_per_launch"],
                  "source": "%s (dram__bytes_read.sum + dram__bytes_write.sum, mean over the %d "
                            "conv_igemm_kernel<1, 5, 0> launches = the CTA-pair 5x5 KPCN layers, forward + data gradient)"
                            % (path, k5["launches"]),
                  "per_kernel": ours}, indent=1))
