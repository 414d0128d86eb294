"""Aggregate an ncu SASS-page export by source line / device function, using nvdisasm line info of the in-tree cubin.

usage: python tools/ncu_by_line.py <report.ncu-rep> <kernel mangled-name substring> [topN]
"""
import csv, io, re, subprocess, sys, collections, os, tempfile

rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
so = os.environ.get("UHC_PROF_SO") or os.path.join(os.path.dirname(__file__), "..", "uhc_b200", "libuhc_b200.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
offs = {}
for cub in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    insec, fn, line = False, "?", ("?", 0)
    for l in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            insec = ksub in m.group(1) and not offs
            fn = "<kernel>"
            continue
        if not insec:
            continue
        m = re.match(r"\s*//## File \"(.*)\", line (\d+)", l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"^(\$\S+|_Z\S+):", l)
        if m and "$" in m.group(1):
            fn = m.group(1).split("$")[-1]
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if m:
            offs[int(m.group(1), 16)] = (fn, line, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ia].startswith("0x")]
base = min(int(r[ia], 16) for r in body)
byfn, byline = collections.Counter(), collections.Counter()
sfn, sline = collections.Counter(), collections.Counter()
stall_fn = collections.defaultdict(collections.Counter)
tot = tots = 0
for r in body:
    o = int(r[ia], 16) - base
    fn, line, _ = offs.get(o, ("?", ("?", 0), ""))
    n, s = int(r[ii]), int(r[isamp])
    tot += n; tots += s
    byfn[fn] += n; byline[line] += n; sfn[fn] += s; sline[line] += s
    for h, i in stall_cols.items():
        stall_fn[fn][h] += int(r[i] or 0)
import shutil
def dem(f):
    try:
        return subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0][-60:]
    except Exception:
        return f
print(f"total warp instructions {tot}  samples {tots}  static SASS {len(body)}")
print("== by device function: inst%  samples%  top stalls")
for fn, n in byfn.most_common():
    st = ", ".join(f"{h[6:]} {100*c/max(1,sfn[fn]):.0f}%" for h, c in stall_fn[fn].most_common(4))
    print(f"{100*n/tot:6.2f}% {100*sfn[fn]/tots:6.2f}%  {dem(fn):60s} {st}")
print("== by source line (inst%, samples%)")
for line, n in byline.most_common(top):
    print(f"{100*n/tot:6.2f}% {100*sline[line]/tots:6.2f}%  {line[0]}:{line[1]}")
if len(sys.argv) > 4:
    key = sys.argv[4]
    col = stall_cols[key]
    bl = collections.Counter(); bf = collections.Counter()
    for r in body:
        o = int(r[ia], 16) - base
        fn, line, _ = offs.get(o, ("?", ("?", 0), ""))
        bl[line] += int(r[col] or 0); bf[fn] += int(r[col] or 0)
    t = sum(bl.values())
    print(f"== {key}: {t} samples ({100*t/tots:.1f}% of all)")
    for line, n in bl.most_common(30):
        print(f"{100*n/t:6.2f}%  {line[0]}:{line[1]}")
