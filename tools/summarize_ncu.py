"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/rNN_launches.txt
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/rNN_kernel.txt
"""
import collections, csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active", "sm__cycles_active.avg",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "sm__maximum_warps_per_active_cycle_pct", "launch__shared_mem_per_block_dynamic",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(src, dst):
    rows = [l for l in open(src) if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(io.StringIO("".join(rows))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        name = row["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# source: {src}; total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
        f.write(f"{'kernel':64s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>10s}\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k[:64]:64s} {v[0]:8d} {v[1]:12.1f} {v[1] / tot:7.3f} {v[1] / v[0]:10.1f}\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source: {src}\n")
        for n, row in enumerate(r[2:]):
            f.write(f"\n## launch {n}: {row[hdr.index('Kernel Name')]}\n")
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write(f"{h:72s} {row[i]:>16s} {units[i]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
