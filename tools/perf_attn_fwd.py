"""Tensor-core attention forward: first kernel (probabilities through shared memory, one CTA per SM) against the second
(probabilities in tensor memory, two CTAs per SM) at the C2 / C5 shapes (CUDA events, warm)."""
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
    res = dict(B=B, S=S, H=H, heads=heads, causal=causal, p=p)
    outs = {}
    for variant in (1, 2):
        ops._lib.call("asme_b200_tc_attn_tune", 2, variant)
        ctx, st, keep = ops.tc_attn_fwd(qkv, kv, B, S, heads, causal, p, 1234, 7, True)
        outs[variant] = (ctx.float(), st, keep)
        res[f"v{variant}_ms"] = round(timeit(lambda: ops.tc_attn_fwd(qkv, kv, B, S, heads, causal, p, 1234, 7, True), iters=20), 4)
    res["max_abs_diff"] = float((outs[1][0] - outs[2][0]).abs().max())
    res["max_abs"] = float(outs[1][0].abs().max())
    res["stats_equal"] = bool(torch.equal(outs[1][1], outs[2][1]))
    res["keep_equal"] = bool(outs[1][2] is None or torch.equal(outs[1][2], outs[2][2]))
    ops._lib.call("asme_b200_tc_attn_tune", 2, 2)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    for cfg in [(256, 200, 64, 2, False, 0.2), (1024, 200, 128, 2, False, 0.0), (1024, 50, 64, 2, True, 0.2), (64, 256, 64, 4, True, 0.1),
                (64, 130, 64, 1, False, 0.0), (128, 40, 64, 2, False, 0.0)]:
        run(*cfg)
