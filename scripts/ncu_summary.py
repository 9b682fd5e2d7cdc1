"""Summarise an .ncu-rep: per-kernel headline metrics + hottest source lines (development aid / profiles/)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
kernel = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "smsp__warps_eligible.avg.per_cycle_active"]
names = [r[hdr.index("Kernel Name")][:34] for r in rows[2:]]
print("%-72s" % "metric", *["%-36s" % n for n in names])
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print("%-72s" % (k[:60] + " [" + units[i] + "]"), *["%-36s" % r[i][:14] for r in rows[2:]])
if kernel:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = None
    lines = {}
    cur_file = None
    for r in rows:
        if r and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        if r and r[0] == "Line No":
            h = r
            continue
        if h and len(r) > 5 and r[0].isdigit():
            try:
                s_ = int(r[h.index("Warp Stall Sampling (All Samples)")])
            except Exception:
                s_ = 0
            key = (cur_file, int(r[0]), r[1][:100])
            lines[key] = lines.get(key, 0) + s_
    tot = sum(lines.values()) or 1
    print("\nhottest source lines of", kernel, "(warp stall samples, first matching launch set)")
    for (f, n, t), s_ in sorted(lines.items(), key=lambda kv: -kv[1])[:28]:
        print("%6.2f%%  %s:%d  %s" % (100.0 * s_ / tot, f, n, t))
