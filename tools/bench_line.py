"""stdin: bench.py output; prints `<tag> ms_per_step value gemm_TFLOP/s sm_MHz loss` (same-box A/B logs under profiles/)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "run"
line = [l for l in sys.stdin if l.startswith("{")]
if not line:
    print(tag, "NO JSON LINE")
    sys.exit(0)
d = json.loads(line[-1])
print(tag, d["ms_per_step"], d["value"], d.get("roofline", {}).get("achieved"), d.get("clocks", {}).get("sm_mhz"), d.get("loss"), flush=True)
