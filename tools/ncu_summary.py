"""Summarise ncu outputs: `launches <csv>` aggregates a gpu__time_duration launch list per kernel;
`raw <ncu-rep>` prints the roofline-relevant counters of every profiled launch."""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows) - 1} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} n={v[0]:5d} total_ms={v[1] / 1e6:9.3f} avg_us={v[1] / v[0] / 1e3:9.2f} share={v[1] / tot:.3f}")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "sm__maximum_warps_per_active_cycle_pct", "launch__waves_per_multiprocessor"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    units = rows[1]
    for r in rows[2:]:
        print("## launch", r[hdr.index("ID")], r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")],
              "block", r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if any(h.startswith(k) for k in KEYS):
                print(f"   {h:75s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
