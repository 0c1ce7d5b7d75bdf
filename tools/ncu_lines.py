"""Per-source-line view of an ncu report (stall samples + executed instructions), with inlined code attributed
to its outermost call site in the kernel's own .cu file.

    python tools/ncu_lines.py REPORT.ncu-rep LAUNCH_INDEX MANGLED_SUBSTRING SOURCE.cu [top_n]

The SASS order of `ncu --page source --csv` is matched with `nvdisasm -gi` of the cubin inside
dgnn_b200/csrc/libdgnn_b200.so (build with -lineinfo)."""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, launch, mangled, srcname = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
top_n = int(sys.argv[5]) if len(sys.argv) > 5 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "dgnn_b200", "csrc", "libdgnn_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, os.path.splitext(srcname)[0] + ".sm_100a.cubin")
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.split("\n")
starts = [i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l]
assert len(starts) == 1, "mangled substring matches %d functions" % len(starts)
end = next(i for i in range(starts[0] + 1, len(dis)) if ".section" in dis[i] and ".text." in dis[i]) \
    if any(".section" in l and ".text." in l for l in dis[starts[0] + 1:]) else len(dis)
insts, pend, cur = [], [], []
for ln in dis[starts[0]:end]:
    if "//##" in ln:
        pend.append(ln.strip()); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        if pend:
            cur, pend = pend, []
        insts.append((m.group(2), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
print(rows[0][:2])
hdr = rows[1]
data = [r for r in rows[2:] if r and r[0].startswith("0x")]
if len(data) == 2 * len(insts):
    data = data[:len(insts)]
assert len(data) == len(insts), (len(data), len(insts))
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
stallcols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
pat = re.compile(re.escape(srcname) + r'", line (\d+)')


INNER = os.environ.get("INNER") == "1"       # attribute to the innermost source line of the kernel's own file


def site(cur):
    best = None
    for c in cur:
        ms = list(pat.finditer(c))
        if INNER and ms:
            return int(ms[0].group(1))
        for m in ms:
            best = int(m.group(1))
    return best


agg = defaultdict(lambda: [0, 0, defaultdict(int)])
for r, (t, cur) in zip(data, insts):
    k = site(cur)
    agg[k][0] += int(r[iI]); agg[k][1] += int(r[iS])
    for c in stallcols:
        v = int(r[c]) if r[c] else 0
        if v:
            agg[k][2][hdr[c][6:]] += v
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
src = open(os.path.join(root, "dgnn_b200", "csrc", srcname)).read().split("\n")
print("warp instructions %d, samples %d" % (tot, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    st = " ".join("%s:%d" % kv for kv in sorted(v[2].items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% samp %5.1f%% inst  L%-4s %-72s %s" % (100 * v[1] / ts, 100 * v[0] / tot, k, src[k - 1].strip()[:72] if k else "", st))
