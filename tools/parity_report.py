"""gpurun_out/parity_records.jsonl (written by the GPU tests, tests/_record.py) -> a readable table.
    python tools/parity_report.py gpurun_out/parity_records.jsonl > profiles/parity_r02.txt"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
# a test that ran twice (a focused pass, then the whole suite) keeps its last record per (test, model)
last_of = {}
for i, r in enumerate(rows):
    last_of[(r["test"], r.get("model"), tuple(sorted(k for k in r if k not in ("test",))))] = i
rows = [r for i, r in enumerate(rows) if i in set(last_of.values())]
print("# measured parity errors of the GPU test run (relative L2 unless named otherwise); bound = what the test asserts")
last = None
for r in rows:
    if r["test"] != last:
        print("\n## %s" % r["test"])
        last = r["test"]
    items = ["%s=%s" % (k, ("%.3e" % v) if isinstance(v, float) else v) for k, v in r.items() if k != "test"]
    print("  " + "  ".join(items))
