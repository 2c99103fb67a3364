#!/usr/bin/env python
"""Summarise an .ncu-rep (from `ncu --set full`) into a small CSV of the metrics DESIGN.md / bench.py quote:
  python profiles/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x_summary.csv"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = [hdr.index(k) for k in KEYS if k in hdr]
out = csv.writer(sys.stdout)
out.writerow([hdr[i] + (" [%s]" % units[i] if units[i] else "") for i in idx])
for r in rows[2:]:
    out.writerow([r[i] for i in idx])
