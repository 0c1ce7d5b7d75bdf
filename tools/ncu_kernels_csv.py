"""ncu report(s) -> the per-kernel evidence table committed under profiles/ (one row per profiled launch) that bench.py
reads for `roofline.traffic`:

    python tools/ncu_kernels_csv.py OUT.csv CELLS TAG=REPORT.ncu-rep [TAG=REPORT.ncu-rep ...]

TAG names the layer shape of the capture (e.g. f128->128); CELLS is the number of cells every launch of the capture
processed (tools/exp_layer_one.py prints it)."""
import csv, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}

out, cells = sys.argv[1], int(sys.argv[2])
rows_out = [["layer", "cells", "Kernel Name"] + METRICS + ["dram_bytes_per_cell"]]
for spec in sys.argv[3:]:
    tag, rep = spec.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.split("\n")))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        vals = []
        for m in METRICS:
            i = hdr.index(m) if m in hdr else -1
            vals.append(r[i] if i >= 0 else "")
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * UNIT.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
        wr = float(r[hdr.index("dram__bytes_write.sum")]) * UNIT.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
        vals[1], vals[2] = "%.0f" % rd, "%.0f" % wr                     # bytes
        rows_out.append([tag, str(cells), r[hdr.index("Kernel Name")]] + vals + ["%.1f" % ((rd + wr) / cells)])
with open(out, "w", newline="") as fh:
    csv.writer(fh).writerows(rows_out)
print("wrote", out, len(rows_out) - 1, "launches")
