"""Compact per-kernel table from an `ncu --set full` report (reads `ncu -i REP --page raw --csv`).
Usage: python tools/ncu_summary.py REP.ncu-rep [--md]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, k, default=float("nan")):
    if k not in col:
        return default
    try:
        return float(r[col[k]].replace(",", ""))
    except ValueError:
        return default


def scaled(r, k):
    """value in base units (bytes / ns)"""
    v = val(r, k)
    u = units[col[k]] if k in col else ""
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e3, "ms": 1e6, "ns": 1, "s": 1e9,
            "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(u, 1)
    return v * mult


STALLS = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print("| kernel | grid x block | regs | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | dram % | L2 MB | occ % | issue % | "
      "tensor % | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("ub2::", "")
    t = scaled(r, "gpu__time_duration.sum")
    rd, wr = scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum")
    l2 = val(r, "lts__t_sectors.sum") * 32
    stalls = sorted(((val(r, s, 0.0), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                     for s in STALLS), reverse=True)[:3]
    tensor = val(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                 val(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"))
    print(f"| {name} | {r[col['launch__grid_size']]} x {r[col['launch__block_size']]} | "
          f"{val(r, 'launch__registers_per_thread'):.0f} | {t / 1e3:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
          f"{(rd + wr) / t:.0f} | {val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {l2 / 1e6:.0f} | "
          f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
          f"{val(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.0f} | {tensor:.0f} | "
          + ", ".join(f"{n} {v:.1f}" for v, n in stalls) + " |")
