#!/bin/bash
O=gpurun_out/r02zz
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print("bench: %.1f patches/s %.3f ms e2e %.1f roofline frac %.3f 720p %.2f ms launches %d clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["denoise_720p"]["ms_per_frame"], d["gpu_launches"], d["clocks"]))
PY
