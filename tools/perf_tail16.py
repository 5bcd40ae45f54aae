"""Folded-bias scoring shape (Kp = 144): 16-column K tail staged with the 32-byte swizzle (knob 6) on / off."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")]
import torch
from asme_b200 import ops
from tools.perf_score import timeit

for R in (1024, 4096):
    V, H, k = 1_000_003, 128, 10
    gen = torch.Generator(device="cuda").manual_seed(0)
    h = torch.randn(R, H, device="cuda", generator=gen)
    w = torch.randn(V, H, device="cuda", generator=gen) * 0.02
    b = torch.randn(V, device="cuda", generator=gen) * 0.01
    tgt = torch.randint(3, V, (R,), device="cuda", generator=gen)
    hb, wb = ops.cast_bf16_ext(h), ops.cast_bf16_ext(w, b)
    fl = 2.0 * R * V * hb.shape[1]
    outs = {}
    for knob in (0, 1):
        ops._lib.call("asme_b200_tc_score_tune", 6, knob)
        outs[knob] = ops.tc_score_topk(hb, wb, None, k, target=tgt)
        ms = timeit(lambda: ops.tc_score_topk(hb, wb, None, k, target=tgt))
        pr = timeit(lambda: ops._lib.call("asme_b200_tc_score_pipeline_probe", ops._p(hb), R, hb.shape[1], ops._p(wb), V, 1, ops._stream()))
        print(json.dumps(dict(R=R, Kp=hb.shape[1], tail16=knob, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1), probe_ld_ms=round(pr, 4))), flush=True)
    assert torch.equal(outs[0]["topk_idx"], outs[1]["topk_idx"]) and torch.equal(outs[0]["topk_val"], outs[1]["topk_val"])
    assert torch.equal(outs[0]["target_score"], outs[1]["target_score"])
    ops._lib.call("asme_b200_tc_score_tune", 6, 1)
