"""Summarise .ncu-rep captures (ncu -i ... --page raw --csv) into the handful of numbers DESIGN.md / bench.py quote.
usage: python profiles/summarize_ncu.py gpurun_out/prof_rows.ncu-rep [...] > profiles/rNN_ncu_full_summary.txt
       python profiles/summarize_ncu.py --traffic-json profiles/r02_ncu_traffic.json "<note>" <algorithmic bytes> file.ncu-rep
(the second form writes the `roofline.traffic` record bench.py reads: DRAM bytes per launch of the first kernel in the file)"""
import csv
import json
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "--traffic-json":
    dst, note, alg, path = sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    r = rows[2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(k):
        return float(r[col[k]].replace(",", "")) * scale.get(units[col[k]], 1.0)
    rec = {"kernel": r[col["Kernel Name"]][:80], "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
           "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"), "algorithmic_bytes": alg,
           "duration_us": float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0),
           "source": path, "note": note}
    json.dump(rec, open(dst, "w"), indent=1)
    print(json.dumps(rec, indent=1))
    sys.exit(0)

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
