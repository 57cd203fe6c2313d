"""Print the headline numbers and the largest kernel classes of a bench.py JSON line (stdin or file)."""
import json, sys
src = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads([l for l in src.strip().splitlines() if l.startswith("{")][-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "n_gpus")}, "e2e", d.get("e2e", {}).get("value"))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "share_of_step") if k in d.get("roofline", {})})
top = int(sys.argv[2]) if len(sys.argv) > 2 else 18
for k in d.get("kernels", [])[:top]:
    print(f"{k['name']:14s} n={k['launches_per_step']:5.1f} ms={k['ms_per_step']:.4f} avg_us={k['avg_ms'] * 1e3:6.1f} tf={k['tflops']:.0f}")
if d.get("sampling"):
    print("sampling", d["sampling"].get("value"), "eager", d["sampling"].get("eager_showers_per_s"))
