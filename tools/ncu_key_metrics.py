#!/usr/bin/env python
"""Extract the handful of ncu metrics quoted in DESIGN.md / bench.py from an .ncu-rep (run where ncu is installed).

    python tools/ncu_key_metrics.py gpurun_out/prof.ncu-rep > profiles/<name>.csv
"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "metric", "unit", "value"])
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        name = d.get("Kernel Name", ("", "?"))[1].split("(")[0]
        for k in KEYS:
            if k in d:
                w.writerow([name, k, d[k][0], d[k][1]])


if __name__ == "__main__":
    main()
