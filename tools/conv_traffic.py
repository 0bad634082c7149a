"""DRAM traffic per launch of the forward / data-gradient convolution kernels from an `ncu --set full` report of
tools/prof_conv.py (second repetition), next to their algorithmic operand bytes -> profiles/conv_traffic.json
(bench.py's roofline.traffic).  Usage: python tools/conv_traffic.py REP.ncu-rep [batch] > profiles/conv_traffic.json"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4
SHAPES = [("inc.3", 512, 64, 0, 64), ("up4.0", 512, 64, 64, 64), ("down1.3", 256, 128, 0, 128),
          ("down2.3", 128, 256, 0, 256), ("up1.0", 64, 512, 512, 512), ("down4.0", 32, 512, 0, 512)]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def scaled(r, k):
    v = float(r[col[k]].replace(",", ""))
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(units[col[k]], 1)


conv = [r for r in data if "conv_" in r[col["Kernel Name"]] and "reduce" not in r[col["Kernel Name"]]]
conv = conv[-18:]   # second repetition: (fwd+stats, dgrad, wgrad) x 6 shapes
launches = []
for i, (name, h, c0, c1, cout) in enumerate(SHAPES):
    for j, what in enumerate(("fwd+stats", "dgrad")):
        r = conv[3 * i + j]
        cin = c0 + c1
        alg = N * h * h * (cin + cout) * 2 + 9 * cin * cout * 2
        launches.append({"layer": name, "pass": what,
                         "kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("ub2::", ""),
                         "dram_bytes": scaled(r, "dram__bytes_read.sum") + scaled(r, "dram__bytes_write.sum"),
                         "algorithmic_bytes": alg})
print(json.dumps({
    "source": f"ncu --set full --clock-control none, tools/prof_conv.py {N}: {len(launches)} forward / data-gradient launches of "
              f"{len(SHAPES)} layer shapes at batch {N} (second repetition)",
    "traffic_bytes_per_launch": sum(l["dram_bytes"] for l in launches) / len(launches),
    "algorithmic_bytes_per_launch": sum(l["algorithmic_bytes"] for l in launches) / len(launches),
    "launches": launches}, indent=1))
