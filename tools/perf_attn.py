"""Tensor-core attention micro-benchmark (CUDA events, warm): forward and both backward variants at the C2 / C5 shapes."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")]
import torch
from asme_b200 import ops
from tools.perf_score import timeit


def run(B, S, H, heads, causal, p):
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = (torch.randn(B * S, 3 * H, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    lens = torch.randint(S // 4, S + 1, (B,), device="cuda", generator=g)
    kv = (torch.arange(S, device="cuda")[None, :] < lens[:, None])
    dctx = (torch.randn(B * S, H, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    ctx, st, keep = ops.tc_attn_fwd(qkv, kv, B, S, heads, causal, p, 1234, 7, True)
    res = dict(B=B, S=S, H=H, heads=heads, causal=causal, p=p)
    res["fwd_ms"] = round(timeit(lambda: ops.tc_attn_fwd(qkv, kv, B, S, heads, causal, p, 1234, 7, True), iters=20), 4)
    outs = {}
    for variant, wgs in ((0, 2), (1, 2), (1, 4)):
        ops._lib.call("asme_b200_tc_attn_tune", 0, variant)
        ops._lib.call("asme_b200_tc_attn_tune", 1, wgs)
        key = variant if wgs == 2 else "1w4"
        outs[key] = ops.tc_attn_bwd(qkv, kv, B, S, heads, causal, ctx, dctx, st, keep, p).float()
        res[f"bwd{key}_ms"] = round(timeit(lambda: ops.tc_attn_bwd(qkv, kv, B, S, heads, causal, ctx, dctx, st, keep, p), iters=20), 4)
    res["w4_equal"] = bool(torch.equal(outs[1], outs["1w4"]))
    ops._lib.call("asme_b200_tc_attn_tune", 1, 4)
    diff = (outs[0] - outs[1]).abs().max().item()
    res["max_abs_diff"], res["max_abs"] = diff, outs[0].abs().max().item()
    ops._lib.call("asme_b200_tc_attn_tune", 0, 1)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    for cfg in [(256, 200, 64, 2, False, 0.2), (256, 200, 64, 2, True, 0.2), (1024, 50, 64, 2, True, 0.2), (256, 200, 128, 2, False, 0.0),
                (64, 256, 64, 4, True, 0.1), (64, 130, 64, 1, False, 0.0)]:
        run(*cfg)
