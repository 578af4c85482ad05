"""configs[3] timing on its own: 1280x720 KPCN denoise, resident / pipelined end-to-end / single-frame latency, fused
and un-fused head.   python tools/frame_bench.py [frames]"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from wcmc_b200 import dropin, lib, ops  # noqa: E402
from wcmc_b200.synth import make_batch  # noqa: E402

dropin.install()
lib.init(0)
from sbmc import KPCN  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
args = types.SimpleNamespace(steps=n)
for fuse in (True, False):
    ops.FUSE_KERNEL_APPLY = fuse
    r = bench.bench_720p(args, KPCN, make_batch)
    print("fused" if fuse else "unfused", {k: v for k, v in r.items() if k not in ("e2e_note", "workload")}, flush=True)
