"""Summarise an .ncu-rep (raw page) into a few lines per launch: python scripts/ncu_summary.py rep [out.md]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_rd_sectors"), ("lts__t_sector_hit_rate.pct", "l2_hit%"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem"), ("sm__cycles_active.avg", "sm_cyc_active"),
        ("sm__cycles_elapsed.max", "sm_cyc_elapsed"), ("smsp__inst_executed.sum", "inst")]
lines = []
for r in rows[2:]:
    parts = []
    for key, name in want:
        if key in idx:
            v = r[idx[key]]
            try:
                v = "%.4g" % float(v)
            except ValueError:
                v = v[:70]
            parts.append("%s=%s%s" % (name, v, (" " + units[idx[key]]) if units[idx[key]] and name not in ("kernel",) else ""))
    lines.append("- " + "; ".join(parts))
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
