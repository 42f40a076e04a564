"""Per-kernel SASS opcode histogram of libapx.so (the evidence that the contraction kernels are tcgen05 / TMA / TMEM code).
    cuobjdump -sass ap_adapter_b200/libapx.so | python tools/sass_hist.py > profiles/sass_opcodes.txt
"""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "UCGABAR_ARV", "ELECT", "MUFU", "HMMA"]
hist = collections.defaultdict(collections.Counter)
cur = None
pat = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = pat.match(line)
    if m and cur:
        hist[cur][m.group(1).split(".")[0]] += 1
names = list(hist)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode histogram of ap_adapter_b200/libapx.so (cuobjdump -sass, sm_100a), one line per kernel.")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA tensor load / store / reduce,")
print("# UBLKCP = cp.async.bulk (1-D TMA, DSMEM pushes), SYNCS = mbarrier ops, UCGABAR_ARV = barrier.cluster.arrive, ELECT = elect.sync,")
print("# MUFU = special-function unit (ex2 / rsqrt / rcp), HMMA = legacy mma.sync (none expected).")
print(f"{'kernel':88s} " + " ".join(f"{k[:8]:>8s}" for k in KEYS) + "   total")
for fn, name in sorted(zip(names, dem), key=lambda t: t[1]):
    c = hist[fn]
    short = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0].replace("void ", "").replace("apx::", "")[:88]
    print(f"{short:88s} " + " ".join(f"{c.get(k, 0):8d}" for k in KEYS) + f"  {sum(c.values()):6d}")
