"""Turn an .ncu-rep (read with `ncu -i ... --page raw --csv`) or a launch-list csv
into the short text summaries committed under profiles/."""
import collections
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        print("----")
        for k in KEYS:
            if k in hdr:
                print(k, row[hdr.index(k)], units[hdr.index(k)])


def launches(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"# {path}: total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v / 1e6:10.3f} ms {c:5d} launches {100 * v / tot:6.2f} %  {k[:110]}")


if __name__ == "__main__":
    (rep if sys.argv[1].endswith(".ncu-rep") else launches)(sys.argv[1])
