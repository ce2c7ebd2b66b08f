"""Key metrics of every kernel in an .ncu-rep (ncu --page raw --csv) as a compact text table: python tools/ncu_summary.py rep [...]"""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"== {rep}: {r[ci['Kernel Name']][:100]}")
        for k in KEYS:
            if k in ci:
                print(f"   {k:75s} {r[ci[k]]:>16s} {units[ci[k]]}")
