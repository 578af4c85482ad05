"""Compact per-launch summary of an .ncu-rep (read here, no GPU needed).
   python tools/ncu_summary.py gpurun_out/x/prof.ncu-rep [more.ncu-rep ...] > profiles/rNN_x.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("time_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"),
    ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("dram_read_MB", "dram__bytes_read.sum"),
    ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pipe_pct", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_hmma_pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("mem_tensor_pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("tc_smem_wavefronts_pct", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("sm_cycles_active", "sm__cycles_active.avg"),
    ("sm_cycles_elapsed", "sm__cycles_elapsed.max"),
    ("inst_executed", "smsp__inst_executed.sum"),
    ("ipc", "sm__inst_executed.avg.per_cycle_elapsed"),
]


def to_num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    u = unit.lower()
    scale = {"gbyte": 1e3, "mbyte": 1.0, "kbyte": 1e-3, "byte": 1e-6}
    if u in scale:
        return round(x * scale[u], 3)
    if u in ("ns", "nsecond"):
        return round(x / 1e3, 2)
    if u in ("ms", "msecond"):
        return round(x * 1e3, 2)
    if u in ("s", "second"):
        return round(x * 1e6, 2)
    return round(x, 3)


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        print("# %s" % path)
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0]
            print("kernel %s  (launch id %s)" % (name, r[hdr.index("ID")]))
            for label, key in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    v = to_num(r[i], units[i])
                    if label == "smem_dyn_KB" and isinstance(v, float) and units[i].lower() == "kbyte":
                        v = round(v * 1e3, 1)
                    print("    %-24s %s" % (label, v))
            print()


if __name__ == "__main__":
    main()
