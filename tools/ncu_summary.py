"""Condense an .ncu-rep (ncu --set full) into a small text table per kernel launch: the metrics DESIGN.md quotes.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_mma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name"), " id", d.get("ID"))
    for k in KEYS:
        if k in d:
            print(f"  {k:85s} {d[k]:>18s} {units[hdr.index(k)]}")
    stalls = [(float(d[h]), h) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h and d[h]]
    for v, h in sorted(stalls, reverse=True)[:8]:
        print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:40s} {v:8.3f} warps per issue-active cycle")
    print()
