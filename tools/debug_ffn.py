"""where does the fused feed-forward kernel differ from the two-GEMM path? (diagnostic)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")]
import torch
from asme_b200 import ops

for (M, H, FF) in [(148 * 128 * 2, 128, 512), (148 * 128 * 3, 128, 128), (148 * 128 * 2, 128, 1024)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    y = torch.randn(M, H, device="cuda", generator=g).bfloat16()
    x = torch.randn(M, H, device="cuda", generator=g)
    w1 = (torch.randn(FF, H, device="cuda", generator=g) * 0.1).bfloat16()
    w2 = (torch.randn(H, FF, device="cuda", generator=g) * 0.1).bfloat16()
    b1 = torch.randn(FF, device="cuda", generator=g) * 0.1
    b2 = torch.randn(H, device="cuda", generator=g) * 0.1
    a = ops.tc_gemm(y, w1, bias=b1, act=1, out_f32=False, out_bf16=True)["bf16"]
    want = ops.tc_gemm(a, w2, bias=b2, residual=x)["f32"]
    for rep in range(3):
        got = ops.tc_ffn_fused(y, w1, b1, w2, b2, x)["f32"]
        torch.cuda.synchronize()
        bad = (got != want)
        nb = int(bad.sum())
        print(f"M={M} H={H} FF={FF} rep{rep}: {nb} bad elements, tiles//148 {torch.unique(bad.any(1).nonzero().flatten() // 128 // 148).tolist()}", flush=True)
        if not nb:
            continue
        used = got - (want - x)          # the residual values the kernel must have added
        rows = bad.any(1).nonzero().flatten()
        kinds = {}
        for r in rows[:400].tolist():
            for u in range(H // 4):
                if not bool(bad[r, 4 * u:4 * u + 4].any()):
                    continue
                vec = used[r, 4 * u:4 * u + 4]
                kind = "?"
                if float(vec.abs().max()) < 1e-5:
                    kind = "zero"
                else:
                    # same row, other 16-byte unit?
                    xr = x[r].view(-1, 4)
                    e = (xr - vec).abs().max(1).values
                    j = int(e.argmin())
                    if float(e[j]) < 1e-4:
                        kind = f"same row, unit offset {j - u}"
                    else:
                        # same unit, other row (of this tile or any)?
                        e2 = (x[:, 4 * u:4 * u + 4] - vec).abs().max(1).values
                        j2 = int(e2.argmin())
                        if float(e2[j2]) < 1e-4:
                            kind = f"other row, offset {j2 - r}"
                        else:
                            e3 = (x.view(-1, 4) - vec).abs().max(1).values
                            j3 = int(e3.argmin())
                            if float(e3[j3]) < 1e-4:
                                rr, uu = divmod(j3, H // 4)
                                kind = f"row offset {rr - r} unit offset {uu - u}"
                kinds[kind] = kinds.get(kind, 0) + 1
        print("   ", sorted(kinds.items(), key=lambda kv: -kv[1])[:12], flush=True)
        r = int(rows[0])
        print("    row", r, "tile row", r % 128, "bad units", [u for u in range(H // 4) if bool(bad[r, 4 * u:4 * u + 4].any())], flush=True)
