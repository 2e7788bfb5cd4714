"""Stall samples of a step-kernel capture aggregated by barrier-delimited segment and by role (execution count).
usage: python profiles/ncu_phases.py <report.ncu-rep>"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None; items = []
for r in rows:
    if "Source" in r and "# Samples" in r: h = r; continue
    if h is None or len(r) < len(h): continue
    items.append(r)
si = h.index('Source'); ns = h.index('# Samples'); ie = h.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
counts = collections.Counter(int(r[ie] or 0) for r in items)
grid = None
by = collections.defaultdict(lambda: [0, 0, collections.Counter()])
seg = 0
for r in items:
    n = int(r[ie] or 0); s = int(r[ns] or 0)
    if 'BAR.SYNC' in r[si]: seg += 1
    k = (seg, n)
    by[k][0] += s; by[k][1] += 1
    for i in stall_cols:
        if r[i] not in ('', '0'): by[k][2][h[i]] += int(r[i])
tot = sum(v[0] for v in by.values())
print("segment, exec-count-per-SASS-line: share of samples, SASS lines, top stalls")
for k, v in sorted(by.items()):
    if v[0] * 200 > tot: print(f"  seg {k[0]:2d} x{k[1]:7d}: {100*v[0]/tot:5.1f}%  {v[1]:5d} lines  {dict(v[2].most_common(3))}")
