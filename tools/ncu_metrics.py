"""Print the metrics the roofline needs from an ncu report (raw page)."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.per_cycle_active", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "lts__t_sectors_srcunit_tex.sum", "launch__grid_size", "launch__block_size"]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if h in want:
            print("  %-75s %-12s %s" % (h, u, v))
