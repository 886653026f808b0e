"""Summarise an .ncu-rep into the handful of metrics the design discusses (run on the CPU box)."""
import csv, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]

UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(path, traffic_json=None, workload=None, kernel_alias=None):
    """Print the summary; with `traffic_json workload [kernel_alias]` also record the capture's DRAM
    bytes (read + write) of the kernel under profiles/ncu_traffic.json[workload][kernel] (what
    bench.py reports as roofline.traffic for the default shapes)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].split("<")[0].split("::")[-1]
        print(f"== {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:85s} {r[i]:>18s} {units[i]}")
        if traffic_json:
            import json, os
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(k)
                tot += float(r[i].replace(",", "")) * UNIT_BYTES[units[i]]
            tab = json.load(open(traffic_json)) if os.path.exists(traffic_json) else {}
            tab.setdefault(workload, {})[kernel_alias or name] = tot
            k = "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"
            if k in hdr:
                tab[workload][(kernel_alias or name) + ":lsu_pipe_pct"] = float(r[hdr.index(k)].replace(",", ""))
            json.dump(tab, open(traffic_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(*sys.argv[1:])
