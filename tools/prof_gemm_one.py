"""ncu target: the four dense layers of the C5 evaluation step, one launch each inside a cudaProfilerStart/Stop region
(`ncu --profile-from-start off ...`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")]
import torch
from asme_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
cases = []
for (M, N, K, gelu, res) in [(204800, 512, 128, 1, 0), (204800, 128, 512, 0, 1), (204800, 384, 128, 0, 0), (204800, 128, 128, 0, 1)]:
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g) if res else None
    cases.append((a, w, dict(bias=bias, act=1 if gelu else 0, residual=r, out_f32=bool(res), out_bf16=not res)))
for _ in range(3):
    for a, w, kw in cases:
        ops.tc_gemm(a, w, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for a, w, kw in cases:
    ops.tc_gemm(a, w, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
