"""Per-kernel counts of the Blackwell-native SASS opcodes in the in-tree libwcmc.so (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA would be the legacy mma.sync path (none expected).
    python tools/sass_opcodes.py > profiles/sass_opcodes.txt        (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "wcmc_b200", "libwcmc.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                       text=True).stdout.split("\n")
OPS = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCCP",
       "SYNCS", "HMMA", "LDGSTS", "ATOMG", "REDG", "RED.")
counts = []
cur = None
i = -1
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        i += 1
        cur = collections.Counter()
        short = re.sub(r"\(.*", "", names[i].replace("(anonymous namespace)::", "")).replace("wcmc::", "")
        counts.append((short, cur))
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cur["_total"] += 1
        for o in OPS:
            if op.startswith(o):
                cur[o] += 1
print("# SASS opcode counts per kernel of %s (cuobjdump -sass; sm_100a)" % os.path.relpath(so, ROOT))
print("# UTC?MMA = tcgen05.mma   LDTM/STTM = tcgen05.ld/st   UTMALDG/UTMASTG = TMA load/store   UTCBAR = tcgen05.commit")
print("%-58s %7s %s" % ("kernel", "instrs", "native opcodes"))
for name, c in sorted(counts):
    nat = "  ".join("%s=%d" % (o, c[o]) for o in OPS if c[o])
    print("%-58s %7d %s" % (name[:58], c["_total"], nat))
tot = collections.Counter()
for _, c in counts:
    tot.update(c)
print("\nTOTAL  " + "  ".join("%s=%d" % (o, tot[o]) for o in OPS if tot[o]))
assert tot["HMMA"] == 0, "legacy mma.sync found"
