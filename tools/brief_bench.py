"""Condensed view of one bench.py JSON line: headline numbers, then the top kernels of the training and evaluation steps."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
print("train", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "launches/step", d.get("gpu_launches_per_step"),
      "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
print("roofline", {k: d["roofline"][k] for k in ("kernel", "bound", "achieved", "frac", "avg_launch_ms", "share_of_step")})
for r in d["kernels"][:n]:
    print(f"  {r['kernel']:22s} {r['shape'][:60]:60s} n={r['launches_per_step']:<4} ms={r['avg_ms']:.4f} share={r['share']:.3f} {r['bound']} frac={r['frac']:.3f}")
e = d.get("eval")
if e:
    print("eval", round(e["value"]), e["unit"], "ms/step", round(e["ms_per_step"], 4), "scoring", e.get("scoring_kernel"))
    for r in e["kernels"][:n]:
        print(f"  {r['kernel']:22s} {r['shape'][:60]:60s} n={r['launches_per_step']:<4} ms={r['avg_ms']:.4f} share={r['share']:.3f} {r['bound']} frac={r['frac']:.3f}")
