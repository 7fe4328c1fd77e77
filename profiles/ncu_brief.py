"""One line per kernel of an ncu report: duration, DRAM bytes, instructions, issue/occupancy."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
def col(r, k):
    return r[hdr.index(k)] if k in hdr else "?"
for r in rows[2:]:
    print(f"{col(r,'Kernel Name')[:44]:44s} {col(r,'gpu__time_duration.sum'):>9s} {rows[1][hdr.index('gpu__time_duration.sum')]:3s} rd {col(r,'dram__bytes_read.sum'):>9s} {rows[1][hdr.index('dram__bytes_read.sum')]:6s}"
          f" wr {col(r,'dram__bytes_write.sum'):>9s} {rows[1][hdr.index('dram__bytes_write.sum')]:6s} inst {float(col(r,'smsp__inst_executed.sum'))/1e6:8.1f}M"
          f" issue {float(col(r,'smsp__issue_active.avg.pct_of_peak_sustained_active')):5.1f}% warps {float(col(r,'sm__warps_active.avg.pct_of_peak_sustained_active')):5.1f}%"
          f" regs {col(r,'launch__registers_per_thread')} dram% {float(col(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):5.1f}")
