"""Key metrics of one-kernel `ncu --set full` reports -> CSV rows (used to build profiles/*.csv).
usage: python scripts/ncu_summary.py report1.ncu-rep [report2 ...] > summary.csv"""
import csv, subprocess, sys

KEYS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
]
w = csv.writer(sys.stdout)
w.writerow(["report", "kernel"] + KEYS)
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        vals = []
        for k in KEYS:
            v = rec.get(k, "")
            vals.append((v + " " + u.get(k, "")).strip())
        w.writerow([rep.split("/")[-1], rec.get("Kernel Name", "")[:60]] + vals)
