"""Summarise an .ncu-rep (ncu --set full) into a small table: one row per kernel launch."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd_MB"),
        ("dram__bytes_write.sum", "dram_wr_MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_inst"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1_ld_sectors"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts")]
idx = [(hdr.index(k) if k in hdr else None, n) for k, n in want]
print(" | ".join(n for _, n in idx))
for r in rows[2:]:
    out = []
    for i, n in idx:
        v = r[i] if i is not None else "NA"
        if n == "kernel":
            v = v.split("(")[0].replace("void uvgb::", "")[:44]
        else:
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
        out.append(v)
    print(" | ".join(out))
