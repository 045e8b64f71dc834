"""SASS evidence per kernel of the shipped library (runs in the build container, no GPU needed).

    python tools/sass_counts.py > profiles/r2_sass_counts.txt

Counts, per kernel of deepcharuco_b200/libdeepcharuco_b200.so: UTCHMMA (tcgen05.mma; `.2CTA` = cta_group::2), UTCBAR (tcgen05.commit),
LDTM (tcgen05.ld), UTMALDG (cp.async.bulk.tensor), SYNCS (mbarrier) and legacy HMMA (mma.sync) instructions.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "deepcharuco_b200", "libdeepcharuco_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names, counts = [], {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        names.append(cur)
        counts[cur] = dict(mma=0, cta2=0, bar=0, ldtm=0, tma=0, syncs=0, hmma=0)
        continue
    if cur is None or "/*" not in line:
        continue
    c = counts[cur]
    if "UTCHMMA" in line:
        c["mma"] += 1
        c["cta2"] += ".2CTA" in line
    elif re.search(r"\bHMMA\b", line):
        c["hmma"] += 1
    if "UTCBAR" in line:
        c["bar"] += 1
    if "LDTM" in line:
        c["ldtm"] += 1
    if "UTMALDG" in line:
        c["tma"] += 1
    if "SYNCS" in line:
        c["syncs"] += 1
dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"dcu::\(anonymous namespace\)::|dcu::<unnamed>::", "", n)
    if n.startswith("conv_tc2_kernel") or n.startswith("conv3x3_tc_kernel"):
        n = n[: n.index(">") + 1].replace("(int)", "").replace("(bool)0", "false").replace("(bool)1", "true")
    return n


print("SASS evidence (cuobjdump -sass deepcharuco_b200/libdeepcharuco_b200.so, built by __graft_entry__.build(); sm_100a only; tools/sass_counts.py).")
print("Per kernel: tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2), tcgen05.commit -> UTCBAR, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG,")
print("mbarrier -> SYNCS.  Legacy HMMA (mma.sync) is counted as a whole word, i.e. not the substring of UTCHMMA.\n")
print("%-110s %8s %6s %7s %5s %8s %6s %12s" % ("kernel", "UTCHMMA", ".2CTA", "UTCBAR", "LDTM", "UTMALDG", "SYNCS", "legacy HMMA"))
tot = dict(mma=0, cta2=0, bar=0, ldtm=0, tma=0, syncs=0, hmma=0)
for n, d in zip(names, dem):
    c = counts[n]
    for k in tot:
        tot[k] += c[k]
    print("%-110s %8d %6d %7d %5d %8d %6d %12d" % (short(d)[:110], c["mma"], c["cta2"], c["bar"], c["ldtm"], c["tma"], c["syncs"], c["hmma"]))
print("%-110s %8d %6d %7d %5d %8d %6d %12d" % ("TOTAL (%d kernels)" % len(names), tot["mma"], tot["cta2"], tot["bar"], tot["ldtm"], tot["tma"], tot["syncs"], tot["hmma"]))
