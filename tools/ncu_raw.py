"""stdin: `ncu --page raw --csv` -> one line per kernel launch with the DRAM bytes / duration / tensor-pipe activity the round summaries quote"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
want = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum")
want += ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
cols = [i for i, h in enumerate(hdr) if h in want]
for r in rows[2:]:
    print(r[ki][:70].ljust(70), "  ".join("%s=%s %s" % (hdr[i].replace("sm__", "").replace("dram__", ""), r[i], units[i]) for i in cols))
