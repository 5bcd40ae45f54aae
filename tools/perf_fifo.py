"""Top-k sweep vs candidate-FIFO depth (knob 7 of asme_b200_tc_score_tune): the FIFOs and the B ring share the CTA's shared
memory, so a shallower FIFO can buy the ring another slot.  Results must not depend on the knob."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")]
import torch
from asme_b200 import ops
from tools.perf_score import timeit

for R, V, H, fold in [(1024, 1_000_003, 128, True), (1024, 1_000_003, 128, False), (4096, 1_000_003, 128, True), (1024, 1_000_003, 64, True)]:
    gen = torch.Generator(device="cuda").manual_seed(0)
    h = torch.randn(R, H, device="cuda", generator=gen)
    w = torch.randn(V, H, device="cuda", generator=gen) * 0.02
    b = torch.randn(V, device="cuda", generator=gen) * 0.01
    tgt = torch.randint(3, V, (R,), device="cuda", generator=gen)
    hb, wb = (ops.cast_bf16_ext(h), ops.cast_bf16_ext(w, b)) if fold else (ops.cast_bf16(h), ops.cast_bf16(w))
    fl = 2.0 * R * V * hb.shape[1]
    ref = None
    for cap in (16, 12, 10, 8):
        ops._lib.call("asme_b200_tc_score_tune", 7, cap)
        out = ops.tc_score_topk(hb, wb, None, 10, target=tgt)
        if ref is None:
            ref = out
        same = all(torch.equal(ref[k], out[k]) for k in ("topk_idx", "topk_val", "target_score"))
        ms = min(timeit(lambda: ops.tc_score_topk(hb, wb, None, 10, target=tgt), iters=20) for _ in range(3))
        print(json.dumps(dict(R=R, Kp=hb.shape[1], fifo=cap, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1), same=same)), flush=True)
ops._lib.call("asme_b200_tc_score_tune", 7, 12)
