#!/usr/bin/env python
"""Per source line (and per file) share of executed warp instructions and stall samples from an ncu report
captured with --import-source on.  Usage: regions_ncu.py report.ncu-rep [min_pct]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, hdr = None, None
inst, samp, text = collections.Counter(), collections.Counter(), {}
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    try:
        inst[(cur, ln)] += int(float(r[ii] or 0)); samp[(cur, ln)] += int(float(r[si] or 0))
    except (ValueError, IndexError): continue
    text[(cur, ln)] = r[1].strip()[:90]
ti, ts = sum(inst.values()), sum(samp.values())
perfile_i, perfile_s = collections.Counter(), collections.Counter()
for k, v in inst.items(): perfile_i[k[0]] += v
for k, v in samp.items(): perfile_s[k[0]] += v
print(f"total warp instr {ti}, samples {ts}")
for f in perfile_i: print(f"{f:28s} instr {100*perfile_i[f]/ti:5.1f}%  samples {100*perfile_s[f]/max(ts,1):5.1f}%")
for k in sorted(inst):
    pi, ps = 100 * inst[k] / ti, 100 * samp[k] / max(ts, 1)
    if pi >= minpct or ps >= minpct:
        print(f"{k[0][:20]:20s}:{k[1]:4d} {pi:5.2f}% | {ps:5.2f}%  {text[k]}")
if len(sys.argv) > 3:   # regions: file:lo-hi,... 
    for spec in sys.argv[3].split(","):
        f, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        si_ = sum(v for k, v in inst.items() if k[0].startswith(f) and lo <= k[1] <= hi)
        ss_ = sum(v for k, v in samp.items() if k[0].startswith(f) and lo <= k[1] <= hi)
        print(f"REGION {spec:34s} instr {100*si_/ti:5.1f}%  samples {100*ss_/max(ts,1):5.1f}%")
