"""Key raw metrics of every kernel in an .ncu-rep (read with `ncu -i ... --page raw --csv`) as aligned text.  usage: ncu_extract.py rep out.txt"""
import csv, io, subprocess, sys
KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w") as f:
    for r in rows[2:]:
        f.write(f"{'Kernel Name':<90} {r[hdr.index('Kernel Name')]}\n")
        for i, h in enumerate(hdr):
            if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
                f.write(f"{h:<90} {units[i]:<12} {r[i]}\n")
        f.write("\n")
print(open(sys.argv[2]).read()[:3000])
