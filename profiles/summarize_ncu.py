"""Summarise .ncu-rep captures (ncu -i ... --page raw --csv) into the handful of numbers DESIGN.md / bench.py quote.
usage: python profiles/summarize_ncu.py gpurun_out/prof_rows.ncu-rep [...] > profiles/rNN_ncu_full_summary.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("==", path)
    for r in rows[2:]:
        print("kernel:", r[col["Kernel Name"]][:60], "| id", r[col["ID"]])
        for k in KEYS:
            if k in col:
                print("   %-70s %s %s" % (k, r[col[k]], units[col[k]]))
