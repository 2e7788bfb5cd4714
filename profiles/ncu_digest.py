"""Digest of one .ncu-rep: headline metrics + the top stall sites of the source view.
usage: python profiles/ncu_digest.py <report.ncu-rep> [n_top]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90])
    for h, v in zip(hdr, r):
        if h in keys or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h):
            try:
                if "issue_stalled" in h and float(v) < 0.15: continue
            except ValueError: pass
            print(f"  {h:90s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
items = []
for r in rows:
    if "Source" in r and "# Samples" in r: h = r; continue
    if h is None or len(r) < len(h): continue
    try: items.append((int(r[h.index("# Samples")]), r[h.index("Address")][-5:], r[h.index("Source")].strip(), int(r[h.index("Instructions Executed")])))
    except ValueError: pass
tot = sum(i[0] for i in items) or 1
print(f"total samples {tot}, SASS lines {len(items)}, warp-instructions {sum(i[3] for i in items)}")
for idx, it in sorted(enumerate(items), key=lambda x: -x[1][0])[:ntop]:
    prev = items[idx - 1][2][:60] if idx else ""
    print(f"  {100*it[0]/tot:5.1f}%  #{idx:5d} {it[2][:70]:70s} x{it[3]:8d}   (prev: {prev})")
