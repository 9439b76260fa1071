"""Development helper: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use in the shipped
library (cuobjdump -sass of libhgru_b200.so).  UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit,
LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk (1-D bulk copy),
SYNCS = mbarrier ops, ACQBULK / ARRIVES variants as printed.
    python monkey-pose_b200/csrc/devtools/sass_opcounts.py > profiles/r02_sass_opcounts.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "libhgru_b200.so")
OPS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "FFMA", "MUFU")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = {}
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1).split(".")[0]
        counts[kern]["_all"] += 1
        if op in OPS:
            counts[kern][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass %s : instruction counts per kernel (static SASS sites, not dynamic counts)" % os.path.basename(LIB))
print("# %-118s %8s %s" % ("kernel", "instrs", " ".join("%8s" % o for o in OPS)))
for (k, c), nm in zip(counts.items(), names):
    if not any(c[o] for o in ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP")) and "--all" not in sys.argv:
        continue
    nm = re.sub(r"hgru::", "", nm)
    nm = re.sub(r"\(.*", "", nm)[:118]
    print("%-120s %8d %s" % (nm, c["_all"], " ".join("%8d" % c[o] for o in OPS)))
    total.update(c)
print("%-120s %8d %s" % ("TOTAL (tensor-core kernels)", total["_all"], " ".join("%8d" % total[o] for o in OPS)))
