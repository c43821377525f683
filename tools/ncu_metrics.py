#!/usr/bin/env python
"""Key metrics per launch from an .ncu-rep (`ncu -i rep --page raw --csv`), as CSV on stdout.
Usage: python tools/ncu_metrics.py gpurun_out/x.ncu-rep > profiles/x_metrics.csv"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(head)}
    cols = [k for k in KEYS if k in idx]
    w = csv.writer(sys.stdout)
    w.writerow(["id", "kernel"] + ["%s [%s]" % (k, units[idx[k]]) for k in cols])
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("<unnamed>::", "").replace("void ", "")[:60]
        w.writerow([r[idx["ID"]], name] + [r[idx[k]] for k in cols])


if __name__ == "__main__":
    main(sys.argv[1])
