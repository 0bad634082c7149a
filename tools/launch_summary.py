"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel share of one training step.
Usage: python tools/launch_summary.py launches.csv [first_kernel_substring]
The step is delimited by consecutive launches of the first kernel of the step (default: the stem conv)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
first = sys.argv[2] if len(sys.argv) > 2 else "pack_weights_multi"
with open(path) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]


def us(r):
    v = float(r["Metric Value"].replace(",", ""))
    return v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v * (1e3 if r["Metric Unit"].startswith("m") else 1)


names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "") for r in rows]
marks = [i for i, n in enumerate(names) if first in n]
if len(marks) < 2:
    raise SystemExit(f"need two launches of '{first}' to delimit a step; found {len(marks)} in {len(rows)} launches")
a, b = marks[-2], marks[-1]
agg = collections.OrderedDict()
tot = 0.0
for r, n in zip(rows[a:b], names[a:b]):
    key = n if n.startswith("ub2::") else "torch: " + n[:60]
    c = agg.setdefault(key, [0, 0.0])
    c[0] += 1
    c[1] += us(r)
    tot += us(r)
ours = sum(v[0] for k, v in agg.items() if k.startswith("ub2::"))
print(f"one step = launches {a}..{b - 1} of {len(rows)} captured: {b - a} kernels ({ours} ub2::, {b - a - ours} torch), "
      f"{tot / 1e3:.3f} ms summed kernel time (ncu: serialised, cold cache)\n")
print("| kernel | launches | us | share |")
print("|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {t:.1f} | {100 * t / tot:.1f}% |")
