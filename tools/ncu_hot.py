"""Top stall sites of a kernel from an .ncu-rep source page (SASS level, needs --import-source / -lineinfo).
   python tools/ncu_hot.py rep.ncu-rep KERNEL_REGEX [N]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name"')
for blk in blocks[1:2]:
    lines = blk.split("\n")
    print("kernel", lines[0][:100])
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    tot = 0
    for r in rows[1:]:
        if len(r) <= isamp or not r[isamp].isdigit():
            continue
        s = int(r[isamp]); tot += s
        data.append((s, r))
    data.sort(key=lambda x: -x[0])
    print("total samples", tot)
    for s, r in data[:n]:
        st = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i]) for i in stall_cols), reverse=True)[:2]
        print("%6d %5.1f%%  %-70s %s" % (s, 100.0 * s / max(tot, 1), r[isrc].strip()[:70], " ".join("%s=%d" % (b, a) for a, b in st if a)))
