"""Stall samples of a kernel capture aggregated by CUDA source line.
The report's SASS rows are in address order; nvdisasm -g of the library's cubin (built from the same sources) gives the
source line of every SASS instruction (innermost inlined frame), so the two are zipped by address offset.
usage: python profiles/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substr> [n_top]"""
import csv, subprocess, sys, io, collections, re, os, tempfile
rep, lib, key = sys.argv[1:4]
ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    on = False; cur = None
    for l in dis.splitlines():
        if l.startswith("//---") and ".text." in l: on = key in l; continue
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
        if m: line_of[int(m.group(1), 16)] = cur
    if line_of: break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None; items = []
for r in rows:
    if "Source" in r and "# Samples" in r: h = r; continue
    if h is None or len(r) < len(h): continue
    items.append(r)
ad = h.index("Address"); ns = h.index("# Samples"); ie = h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
base = int(items[0][ad], 16)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in items:
    k = line_of.get(int(r[ad], 16) - base)
    agg[k][0] += int(r[ns] or 0); agg[k][1] += int(r[ie] or 0)
    for i in stall_cols:
        if r[i] not in ("", "0"): agg[k][2][h[i][6:]] += int(r[i])
tot = sum(v[0] for v in agg.values()) or 1; toti = sum(v[1] for v in agg.values()) or 1
srcs = {}
def text(k):
    if not k: return ""
    if k[0] not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", k[0])
        srcs[k[0]] = open(p).read().splitlines() if os.path.exists(p) else []
    L = srcs[k[0]]
    return L[k[1] - 1].strip()[:70] if 0 < k[1] <= len(L) else ""
print(f"samples {tot}, warp-instructions {toti}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:ntop]:
    print(f"{100*v[0]/tot:5.1f}% smp {100*v[1]/toti:5.1f}% ins  {str(k):34s} {dict(v[2].most_common(2))}  | {text(k)}")
