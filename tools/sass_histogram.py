"""SASS opcode evidence for the in-tree library: whole-library histogram and, per kernel, the counts of the opcodes that
prove the Blackwell paths (tcgen05 MMA = UTCHMMA, TMEM loads = LDTM, TMA tensor loads / stores = UTMALDG / UTMASTG,
bulk copies = UBLKCP, cp.async = LDGSTS, mbarrier = SYNCS).

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "dgnn_b200", "csrc", "libdgnn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "FFMA", "LDG", "STG", "LDS", "STS",
       "SHFL", "ATOMG", "REDG"]
total = collections.Counter()
per = collections.OrderedDict()
cur = None
for ln in out.split("\n"):
    m = re.match(r"\s+Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("dgnn::", "")
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and cur is not None:
        total[m.group(1)] += 1
        per[cur][m.group(1)] += 1
print("# SASS opcode histogram of dgnn_b200/csrc/libdgnn_b200.so (cuobjdump -sass, sm_100a), tools/sass_histogram.py")
print("# whole library: %d instructions in %d kernels" % (sum(total.values()), len(per)))
for op, n in total.most_common():
    print("%8d  %s" % (n, op))
print()
print("# per kernel: " + " ".join(KEY))
for k, c in per.items():
    if not any(c[o] for o in ("UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS")):
        continue
    print("%-44s %s" % (k[:44], " ".join("%s=%d" % (o, c[o]) for o in KEY if c[o])))
