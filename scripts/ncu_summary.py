"""Curated summary of an .ncu-rep (read on the CPU box): usage: ncu_summary.py rep [out.md] [title]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__inst_executed_pipe_lsu.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# {sys.argv[3] if len(sys.argv) > 3 else rep}", "", f"source: `ncu --set full --clock-control none --import-source on`, report {rep}", ""]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        out.append(f"## {r[name_i]}")
        out.append("")
        out.append("| metric | unit | value |")
        out.append("|---|---|---:|")
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k:
                    out.append(f"| {h} | {units[i]} | {r[i]} |")
        for i, h in enumerate(hdr):
            if "tensor" in h and h.endswith("pct_of_peak_sustained_elapsed") and "ops_path" not in h and h not in KEYS:
                out.append(f"| {h} | {units[i]} | {r[i]} |")
        out.append("")
    txt = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
