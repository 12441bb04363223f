"""Summarise `ncu --metrics dram__bytes_*` over the GEMM launches of a step -> profiles/<tag>_gemm_traffic.json."""
import collections, csv, json, re, sys
path, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if not l.startswith("==")]
per = collections.defaultdict(dict)
for row in csv.DictReader(lines):
    try:
        per[row["ID"]][row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
        per[row["ID"]]["unit:" + row["Metric Name"]] = row["Metric Unit"]
        per[row["ID"]]["kernel"] = re.sub(r"\(.*", "", row["Kernel Name"])
    except Exception:
        pass
def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
tot_b = tot_t = 0.0
tens = []
for k, d in per.items():
    rd = to_bytes(d.get("dram__bytes_read.sum", 0), d.get("unit:dram__bytes_read.sum", "byte"))
    wr = to_bytes(d.get("dram__bytes_write.sum", 0), d.get("unit:dram__bytes_write.sum", "byte"))
    tot_b += rd + wr
    tu = d.get("unit:gpu__time_duration.sum", "ns")
    tot_t += d.get("gpu__time_duration.sum", 0) * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9}.get(tu, 1e-9)
    if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in d:
        tens.append((d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"], d.get("gpu__time_duration.sum", 0)))
n = max(len(per), 1)
wt = sum(t for _, t in tens) or 1.0
res = {"launches": len(per), "dram_bytes_per_launch": round(tot_b / n), "dram_bytes_total": round(tot_b), "time_s_total_under_ncu": tot_t,
       "tensor_pipe_active_pct_time_weighted": round(sum(p * t for p, t in tens) / wt, 2) if tens else None,
       "source": path}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
