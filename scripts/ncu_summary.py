"""Summarise ncu outputs into small text files for profiles/.
  python scripts/ncu_summary.py launches <launches.csv>          -> per-kernel count / mean / share
  python scripts/ncu_summary.py full <report.ncu-rep> [kernel#]   -> key metrics of one captured launch
"""
import csv, subprocess, sys, collections, io

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        val = float(r[-1].replace(",", ""))
        unit = r[-2]
        if unit == "ns": val /= 1e3
        elif unit == "ms": val *= 1e3
        elif unit == "s": val *= 1e6
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += val
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,mean_us,total_us,share")
    for k, (c, t) in agg.items():
        print("%s,%d,%.1f,%.1f,%.3f" % (k, c, t / c, t, t / tot))

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]

def full(path, idx=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    vals = rows[2 + idx]
    d = dict(zip(hdr, zip(units, vals)))
    print("kernel:", d.get("Kernel Name", ("", ""))[1])
    for k in KEYS:
        if k in d:
            print("%-90s %-14s %s" % (k, d[k][0], d[k][1]))

if __name__ == "__main__":
    if sys.argv[1] == "launches": launches(sys.argv[2])
    else: full(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
