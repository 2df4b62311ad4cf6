"""Print a bench breakdown JSON (bench.py --breakdown) grouped by pass and sorted by time; optional second file = diff base."""
import json
import sys

d = json.load(open(sys.argv[1]))
base = json.load(open(sys.argv[2]))["ops"] if len(sys.argv) > 2 else {}
print(f"ms_per_step {d['ms_per_step']:.1f}  instrumented {d['instrumented_ms']:.1f}")
grp = {}
for k, v in d["ops"].items():
    grp[k.split(":")[0]] = grp.get(k.split(":")[0], 0) + v["ms"]
print({k: round(v, 1) for k, v in grp.items()})
for k, v in sorted(d["ops"].items(), key=lambda kv: -kv[1]["ms"]):
    if v["ms"] > 0.7:
        b = base.get(k)
        extra = f"  (was {b['ms']:7.2f})" if b else ""
        print(f"  {k:36s} {v['ms']:8.2f} ms {v['tflops']:7.0f} TF exec {v['tflops_executed']:7.0f} n={v['launches']}{extra}")
