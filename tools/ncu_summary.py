"""tools/ncu_summary.py REPORT.ncu-rep -- key raw metrics of every profiled launch, as text (for profiles/)."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name", "?")[:100])
    for k in WANT:
        if k in d:
            print("  %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
    rd, wr, t = float(d.get("dram__bytes_read.sum", 0)), float(d.get("dram__bytes_write.sum", 0)), float(d.get("gpu__time_duration.sum", 1))
    ur, ut = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("gpu__time_duration.sum")]
    print("  dram traffic (read+write): %.3f %s in %.3f %s" % (rd + wr, ur, t, ut))
