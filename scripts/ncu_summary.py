"""Summarise an .ncu-rep (one kernel launch per id) into the few numbers the roofline argument needs.
usage: python scripts/ncu_summary.py report.ncu-rep [more.ncu-rep ...] > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_no_instructions",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_sleeping",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_sample_buffer_full", "smsp__pcsamp_warps_issue_stalled_tex_throttle",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
            print(f"# {path}")
            print(f"kernel: {d.get('Kernel Name', ('?',))[0]}")
            for k in WANT:
                if k in d and d[k][0] != "":
                    print(f"  {k:75s} {d[k][0]:>16s} {d[k][1]}")
            tens = [(h, d[h]) for h in hdr if "tensor" in h and h.endswith("pct_of_peak_sustained_elapsed") and "avg" in h
                    and d[h][0] not in ("", "0")]
            for h, (v, u) in tens:
                if h not in WANT:
                    print(f"  {h:75s} {v:>16s} {u}")
            print()


if __name__ == "__main__":
    main()
