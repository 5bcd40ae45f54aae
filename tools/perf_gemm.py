"""Tensor-core dense-layer micro-benchmark (CUDA events, warm): both tall-kernel variants at the C2 / C5 shapes."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")]
import torch
from asme_b200 import ops
from tools.perf_score import timeit


def run(M, N, K, kn, gelu, res, drop):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(K, N, device="cuda", generator=g) * 0.1).bfloat16() if kn else (torch.randn(N, K, device="cuda", generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g) if res else None
    kw = dict(b_is_kn=bool(kn), bias=bias, act=1 if gelu else 0, residual=r, p_drop=0.2 if drop else 0.0, seed=11, site=5,
              out_f32=bool(res), out_bf16=not res)
    out = {}
    res_d = dict(M=M, N=N, K=K, kn=kn, gelu=gelu, res=res, drop=drop)
    for variant in (0, 1):
        ops._lib.call("asme_b200_tc_gemm_tune", 0, variant)
        o = ops.tc_gemm(a, w, **kw)
        out[variant] = (o["f32"] if res else o["bf16"]).float()
        res_d[f"v{variant}_ms"] = round(timeit(lambda: ops.tc_gemm(a, w, **kw), iters=20), 4)
    res_d["equal"] = bool(torch.equal(out[0], out[1]))
    nbytes = 2 * (M * K + N * K) + M * N * (4 + 4 if res else 2)
    res_d["v1_gbs"] = round(nbytes / res_d["v1_ms"] / 1e6, 1)
    ops._lib.call("asme_b200_tc_gemm_tune", 0, 1)
    print(json.dumps(res_d), flush=True)


if __name__ == "__main__":
    for cfg in [(51200, 256, 64, 0, 1, 0, 1), (51200, 64, 256, 0, 0, 1, 1), (51200, 192, 64, 0, 0, 0, 0), (51200, 64, 64, 0, 0, 1, 1),
                (51200, 64, 256, 1, 0, 0, 0), (51200, 256, 64, 1, 0, 0, 0),
                (204800, 512, 128, 0, 1, 0, 0), (204800, 128, 512, 0, 0, 1, 0), (204800, 384, 128, 0, 0, 0, 0), (204800, 128, 128, 0, 0, 1, 0),
                (1024, 128, 128, 0, 0, 1, 0), (100, 96, 64, 0, 0, 0, 0)]:
        run(*cfg)
