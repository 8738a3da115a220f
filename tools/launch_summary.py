"""Summarises an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list of
`bench.py --no-graph`: per-kernel launch count, total time, share and (when captured) DRAM bytes over ONE steady-state training
step (the launches between the last two optimizer kernels).  With a second argument it also writes the JSON bench.py reads for
`roofline.traffic` (average dram bytes per tcgen05 GEMM launch of that step).
  python tools/launch_summary.py gpurun_out/launches_r02k.csv profiles/gemm_traffic_r02.json > profiles/launches_r02k_summary.txt"""
import collections
import csv
import json
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
iid, iname, imet, ival = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
launches = collections.OrderedDict()     # id -> {name, metrics}
for x in r:
    if len(x) <= ival:
        continue
    e = launches.setdefault(x[iid], {"name": x[iname]})
    e[x[imet]] = float(x[ival].replace(",", ""))
data = list(launches.values())
opt = [i for i, e in enumerate(data) if "sgd_kernel" in e["name"] or "adamw_kernel" in e["name"]]
seg = data[opt[-2] + 1:opt[-1] + 1] if len(opt) >= 2 else data


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"\(.*$", "", n)
    return n[:72]


T = "gpu__time_duration.sum"
has_dram = any("dram__bytes_read.sum" in e for e in seg)
agg = collections.OrderedDict()
for e in seg:
    k = short(e["name"])
    c, t, b = agg.get(k, (0, 0.0, 0.0))
    agg[k] = (c + 1, t + e.get(T, 0.0), b + e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0))
tot = sum(t for _, t, _ in agg.values())
print("%s: one steady-state step = %d launches, %.3f ms of kernel time (cold-cache, serialised: compare shares)" % (path, len(seg), tot / 1e6))
for k, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    extra = "  %8.1f MB DRAM (%.0f GB/s)" % (b / 1e6, b / max(t, 1.0)) if has_dram else ""
    print("%-74s %4d %9.3f ms %5.1f%%%s" % (k, c, t / 1e6, 100.0 * t / tot, extra))
if len(sys.argv) > 2 and has_dram:
    g = [e for e in seg if "vitb_gemm_kernel" in e["name"] or "vitb_wgrad_pair_kernel" in e["name"]]
    byt = sum(e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0) for e in g)
    out = {"source": path, "gemm_launches_per_step": len(g), "dram_bytes_per_launch": byt / max(1, len(g)),
           "dram_bytes_per_step": byt, "gemm_ms_per_step_under_ncu": sum(e.get(T, 0.0) for e in g) / 1e6,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum over the tcgen05 GEMM launches of one steady-state step, ncu "
                   "--clock-control none, per-launch values are cold-cache and serialised"}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print("wrote %s: %d GEMM launches, %.1f MB per launch" % (sys.argv[2], len(g), byt / max(1, len(g)) / 1e6))
