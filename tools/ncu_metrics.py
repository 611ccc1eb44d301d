"""Key metrics per kernel instance from an .ncu-rep (via `ncu -i rep --page raw --csv`)."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("### " + r[idx["Kernel Name"]][:110])
    for w in want:
        if w in idx:
            print("  %-68s %s %s" % (w, r[idx[w]][:40], units[idx[w]]))
