"""Per-kernel table of ONE training step from an ncu CSV with the metrics
    gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum, sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
-> launches, total/avg time, DRAM bytes per launch, achieved DRAM GB/s and its fraction of the measured HBM peak (MEASURED_PEAKS.json
hbm_gbs, default 6570.6), time-weighted tensor-pipe active %.  Also writes <out>.json with the same rows plus the GEMM-family aggregate
(`dram_bytes_per_launch` = roofline.traffic of bench.py).

    python tools/step_kernel_table.py gpurun_out/r02_step_metrics.csv profiles/r02_step_kernels   ->  .txt + .json
"""
import collections
import csv
import json
import os
import re
import sys

path, out = sys.argv[1], sys.argv[2]
peak = 6570.6
mp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(mp):
    peak = json.load(open(mp)).get("hbm_gbs", peak)
lines = [l for l in open(path) if not l.startswith("==")]
per = collections.OrderedDict()
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3,
         "s": 1.0, "second": 1.0, "%": 1.0}
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    d = per.setdefault(row["ID"], {"kernel": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")})
    d[row["Metric Name"]] = v * SCALE.get(row["Metric Unit"], 1)
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["kernel"], dict(launches=0, time_s=0.0, dram_bytes=0.0, tensor_w=0.0))
    t = d.get("gpu__time_duration.sum", 0.0)
    a["launches"] += 1
    a["time_s"] += t
    a["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["tensor_w"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t
tot = sum(a["time_s"] for a in agg.values()) or 1.0
rows = []
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_s"]):
    gbs = a["dram_bytes"] / a["time_s"] / 1e9 if a["time_s"] > 0 else 0.0
    rows.append(dict(kernel=k, share=a["time_s"] / tot, launches=a["launches"], total_ms=a["time_s"] * 1e3, avg_us=a["time_s"] / a["launches"] * 1e6,
                     dram_mb_per_launch=a["dram_bytes"] / a["launches"] / 1e6, dram_gbs=gbs, hbm_frac=gbs / peak,
                     tensor_pipe_pct=a["tensor_w"] / a["time_s"] if a["time_s"] > 0 else 0.0))
gemm = [r for r in rows if "gemm_tcgen05" in r["kernel"] or "gemm_thin_cluster" in r["kernel"]]
gl = sum(r["launches"] for r in gemm) or 1
gt = sum(r["total_ms"] for r in gemm) or 1.0
summary = dict(source=path, hbm_peak_gbs=peak, launches=len(per), kernel_time_ms_under_ncu=tot * 1e3,
               gemm_family=dict(launches=gl, total_ms=gt, dram_bytes_per_launch=round(sum(r["dram_mb_per_launch"] * r["launches"] for r in gemm) * 1e6 / gl),
                                tensor_pipe_active_pct_time_weighted=round(sum(r["tensor_pipe_pct"] * r["total_ms"] for r in gemm) / gt, 2)),
               kernels=rows)
json.dump(summary, open(out + ".json", "w"), indent=1)
with open(out + ".txt", "w") as f:
    f.write(f"# one eager C3 step under ncu (--clock-control none; per-launch times are cold-cache and serialised): {len(per)} launches, {tot*1e3:.2f} ms of kernel time\n")
    f.write(f"# DRAM GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum; HBM peak {peak} GB/s (MEASURED_PEAKS.json)\n")
    f.write(f"{'share':>7} {'total_ms':>9} {'n':>5} {'avg_us':>8} {'MB/launch':>10} {'GB/s':>8} {'of_HBM':>7} {'tensor%':>8}  kernel\n")
    for r in rows:
        f.write(f"{r['share']*100:6.2f}% {r['total_ms']:9.3f} {r['launches']:5d} {r['avg_us']:8.1f} {r['dram_mb_per_launch']:10.2f} {r['dram_gbs']:8.0f} "
                f"{r['hbm_frac']*100:6.1f}% {r['tensor_pipe_pct']:8.1f}  {r['kernel'][:90]}\n")
print(open(out + ".txt").read())
