"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --no-graph`: per-kernel launch count,
total time and share over ONE steady-state training step (the launches between the last two optimizer kernels).
  python tools/launch_summary.py gpurun_out/launches_r02b.csv > profiles/launches_r02b_summary.txt"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
iname, ival = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [(x[iname], float(x[ival].replace(",", ""))) for x in r if len(x) > ival]
opt = [i for i, (n, _) in enumerate(data) if "sgd_kernel" in n or "adamw_kernel" in n]
if len(opt) >= 2:
    seg = data[opt[-2] + 1:opt[-1] + 1]
else:
    seg = data


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"\(.*$", "", n)
    return n[:72]


agg = collections.OrderedDict()
for n, v in seg:
    k = short(n)
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + v)
tot = sum(t for _, t in agg.values())
print("%s: one steady-state step = %d launches, %.3f ms of kernel time (cold-cache, serialised: compare shares)" % (path, len(seg), tot / 1e6))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-74s %4d %9.3f ms %5.1f%%" % (k, c, t / 1e6, 100.0 * t / tot))
