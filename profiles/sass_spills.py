"""List the local-memory (spill) instructions of one kernel's SASS with the nearest preceding CALL/BAR, so that
spills on the hot path can be told from spills around cold calls.  usage: sass_spills.py <lib.so> <kernel-substr>"""
import subprocess, sys, re
lib, key = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
on = False; lines = []
for l in out.splitlines():
    if "Function :" in l: on = key in l
    elif on and re.search(r"/\*[0-9a-f]{4,5}\*/", l): lines.append(l.strip())
n = 0
for i, l in enumerate(lines):
    if re.search(r"\b(STL|LDL)\b", l):
        n += 1
        ctx = [x for x in lines[max(0, i - 40):i] if re.search(r"\b(CALL|BAR|BRA|EXIT|RET)\b", x)][-1:]
        print(i, l[:80], " | after:", ctx[0][10:60] if ctx else "")
print("total", n, "of", len(lines))
