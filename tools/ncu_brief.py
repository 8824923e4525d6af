"""Compact view of one `ncu --set full` report (first kernel): duration, pipes, memory, top stalls."""
import csv
import subprocess
import sys


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
    d = dict(zip(hdr, vals))

    def g(k):
        return d.get(k, "?")
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
    units = dict(zip(hdr, rows[1])) if len(rows) > 2 else {}
    for k in keys:
        print(f"{k:70s} {g(k)} {units.get(k, '')}")
    st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
          for h, v in d.items() if "smsp__average_warps_issue_stalled" in h and "_per_issue_active" in h]

    def f(x):
        try:
            return float(x.replace(",", ""))
        except Exception:
            return 0.0
    print("stalls (warps per issue):", ", ".join(f"{h}={f(v):.2f}" for h, v in sorted(st, key=lambda hv: -f(hv[1]))[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
