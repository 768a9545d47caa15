"""Turns an .ncu-rep into the short text summary committed under profiles/ (run here, no GPU needed)."""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct",
        "smsp__inst_executed.sum ", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "launch__registers_per_thread ",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct_of_peak_sustained_elapsed")


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[h.index("Kernel Name")])
        for i, n in enumerate(h):
            if any((n + " ").startswith(k) for k in KEEP):
                if "stalled" in n:
                    try:
                        if float(r[i].replace(",", "")) < 0.05:
                            continue
                    except ValueError:
                        pass
                print(f"  {n} [{u[i]}] = {r[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
