"""Condense an .ncu-rep (ncu --set full) into the per-kernel summary format kept under profiles/.
usage: ncu_summary.py report.ncu-rep "header comment" > profiles/rNN_xxx_ncu_full_summary.csv"""
import csv
import io
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__cluster_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum")
STALL = "smsp__average_warps_issue_stalled_"

rep, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# {note}")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"## {d.get('Kernel Name', '?')[:90]} grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
    print("metric,unit,value")
    for k, u in zip(hdr, units):
        if k in KEEP or (k.startswith(STALL) and k.endswith("_per_issue_active.ratio")):
            print(f"{k},{u},{d[k]}")
