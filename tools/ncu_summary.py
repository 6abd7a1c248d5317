"""Summarise an .ncu-rep (read here, no GPU needed): key metrics per kernel launch.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [substring filters...]"""
import csv
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit", "launch__waves_per_multiprocessor",
           "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct",
           "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "smsp__inst_executed.sum ",
           "smsp__average_warps_issue_stalled", "smsp__warp_issue_stalled", "l1tex__data_bank_conflicts",
           "lts__t_bytes.sum ", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_"]


def main():
    rep = sys.argv[1]
    filt = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"=== {name}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for h, u, v in zip(hdr, units, r):
            if any(f.strip() in h for f in filt):
                try:
                    if float(v.replace(",", "")) == 0.0 and "stalled" in h:
                        continue
                except ValueError:
                    pass
                print(f"  {h:90s} {v} {u}")


if __name__ == "__main__":
    main()
