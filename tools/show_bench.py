import json, sys
d = json.load(open(sys.argv[1]))
for k, v in d.items():
    if k not in ('kernels',): print(k, v)
print()
for r in d.get('kernels', []):
    print(f"{r['kernel']:24s} {r['shape']:70s} n/step={r['launches_per_step']:<5} avg_ms={r['avg_ms']:.4f} share={r['share']:.3f} {r['bound']} frac={r['frac']:.3f}")
