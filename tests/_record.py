"""Measured parity errors of the GPU tests, appended to gpurun_out/parity_records.jsonl when that directory exists
(the GPU box); tools/parity_report.py turns the records into profiles/parity_rNN.txt."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(test, **values):
    out = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out):
        return
    with open(os.path.join(out, "parity_records.jsonl"), "a") as f:
        f.write(json.dumps({"test": test, **{k: (float(v) if hasattr(v, "__float__") else v) for k, v in values.items()}}) + "\n")
