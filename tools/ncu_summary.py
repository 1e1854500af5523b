"""Summarise an .ncu-rep (read here, no GPU needed) as a markdown table per kernel launch.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls] > profiles/....md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_blocks",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    path = sys.argv[1]
    want_stalls = "--stalls" in sys.argv
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print("## %s  (launch id %s, grid %s, block %s)\n" % (name[:110], r[col["ID"]], r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
        print("| metric | unit | value |\n|---|---|---|")
        for k in KEYS:
            if k in col:
                print("| %s | %s | %s |" % (k, units[col[k]], r[col[k]]))
        if want_stalls:
            st = [(float(r[i] or 0), h[len(STALLS):].replace("_per_issue_active.ratio", "")) for h, i in col.items()
                  if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")]
            st.sort(reverse=True)
            print("| stalls per issue (top) |  | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in st[:6]))
        print()


if __name__ == "__main__":
    main()
