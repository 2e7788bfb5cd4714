#!/usr/bin/env python
"""Turn an `ncu --set full --import-source on` report into a short text summary (key metrics, stall
reasons per issued instruction, hottest source lines).  Usage: summarise_ncu.py report.ncu-rep [n_lines]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.max"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {rep}")
    for r in data[:1]:
        d = dict(zip(hdr, r))
        print(f"kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"  {h:72s} {r[i]:>16s} {units[i]}")
        print("  stall cycles per issued instruction:")
        st = [(float(r[i]), h) for i, h in enumerate(hdr)
              if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i] not in ("", "n/a")]
        for v, h in sorted(st, reverse=True)[:8]:
            print(f"    {h.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {v:6.2f}")
    rows = ncu_csv(rep, "source", ["--print-source", "cuda,sass"])
    agg, cur, h2, nsec, nwarps = collections.OrderedDict(), None, None, 0, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            nsec += 1
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            h2 = r
            continue
        if h2 is None or r[0] in ("", "Function Name"):
            continue
        try:
            ln = int(r[0])
            ie = float(r[h2.index("Instructions Executed")])
            st = float(r[h2.index("Warp Stall Sampling (All Samples)")])
        except ValueError:
            continue
        o = agg.get((cur, ln), (0, 0, ""))
        agg[(cur, ln)] = (o[0] + ie, o[1] + st, r[1].strip()[:96])
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print(f"  warp-level instructions executed (all sections): {tot:.0f}")
    print("  hottest source lines (share of executed instructions | share of stall samples):")
    for (f, ln), (ie, st, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:nlines]:
        print(f"    {f[:22]:22s}:{ln:4d} {ie / tot * 100:5.1f}% | {st / tots * 100:5.1f}%  {src}")


if __name__ == "__main__":
    main()
