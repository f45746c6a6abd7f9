"""Per-kernel SASS evidence that the library is Blackwell-native: counts of UTCHMMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor loads / stores), UBLKCP (bulk copies), MUFU, HMMA (legacy mma.sync —
expected 0) for every kernel in libimagd_b200.so. Runs without a GPU:
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "imagdressing_b200", "libimagd_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
MN = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU", "HMMA")
kern = None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[kern]["total"] += 1
        for k in MN:
            if op.startswith(k) and not (k == "HMMA" and op.startswith("HMMA") is False):
                counts[kern][k] += 1
print(f"# SASS summary of {os.path.relpath(so, ROOT)} ({os.path.getsize(so)} bytes), sm_100a; HMMA = legacy mma.sync (expected 0)")
print(f"{'kernel':110s} " + " ".join(f"{k:>8s}" for k in ("total",) + MN))
tot = collections.Counter()
for k, c in counts.items():
    name = demangle(k)
    name = re.sub(r"\(CUtensorMap_st.*", "(...)", name)[:110]
    print(f"{name:110s} " + " ".join(f"{c[m]:8d}" for m in ("total",) + MN))
    tot.update(c)
print(f"{'ALL KERNELS (' + str(len(counts)) + ')':110s} " + " ".join(f"{tot[m]:8d}" for m in ("total",) + MN))
