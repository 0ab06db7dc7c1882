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


def traffic(path, key, out="profiles/ncu_traffic.json", pick=None):
    """dram__bytes_read + dram__bytes_write per launch of the kernels in an .ncu-rep (mean over the
    captured launches, or only those whose name contains ``pick``) -> profiles/ncu_traffic.json[key];
    bench.py reports it as roofline.traffic together with the library build it belongs to."""
    import json
    import pathlib
    root = pathlib.Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    from phamclust_b200 import build_native
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    vals, names = [], []
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")]
        if pick and pick not in name:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            k = hdr.index(m)
            tot += float(row[k].replace(",", "")) * scale[units[k]]
        vals.append(tot)
        names.append(name)
    target = root / out
    data = json.loads(target.read_text()) if target.exists() else {}
    data[key] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals),
                 "kernels": sorted(set(n.split("(")[0] for n in names))[:8],
                 "source": f"ncu --set full --clock-control none, {pathlib.Path(path).name}; library sha256 "
                           f"{build_native.library_sha256()[:16]}"}
    target.write_text(json.dumps(data, indent=1) + "\n")
    print(key, data[key])


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], pick=sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        (rep if sys.argv[1].endswith(".ncu-rep") else launches)(sys.argv[1])
