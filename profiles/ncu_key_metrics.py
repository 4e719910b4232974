"""Key `ncu --set full` metrics per kernel launch of a report, as a markdown table:
    python profiles/ncu_key_metrics.py report.ncu-rep > summary.md"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
cols = [(hdr.index(k), n) for k, n in KEYS if k in hdr]
print("| kernel | " + " | ".join(n for _, n in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    print("| `%s` | " % name[:70] + " | ".join("%s %s" % (r[i][:9], units[i]) for i, _ in cols) + " |")
