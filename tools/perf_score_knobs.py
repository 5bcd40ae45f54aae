"""Scoring-kernel knob sweep (CUDA events, warm): CTA pairs x epilogue warpgroups x threshold pass x reject-all, per shape.
    python tools/perf_score_knobs.py
"""
import os, sys, json, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")]
import torch
from asme_b200 import ops
from tools.perf_score import timeit


def tune(knob, value):
    ops._lib.call("asme_b200_tc_score_tune", knob, value)


def run(R, V, H, k, fold):
    gen = torch.Generator(device="cuda").manual_seed(0)
    h = torch.randn(R, H, device="cuda", generator=gen)
    w = torch.randn(V, H, device="cuda", generator=gen) * 0.02
    b = torch.randn(V, device="cuda", generator=gen) * 0.01
    tgt = torch.randint(3, V, (R,), device="cuda", generator=gen)
    if fold:
        hb, wb = ops.cast_bf16_ext(h), ops.cast_bf16_ext(w, b)
    else:
        hb, wb = ops.cast_bf16(h), ops.cast_bf16(w)
    fl = 2.0 * R * V * hb.shape[1]
    ref = None
    configs = [(0, 4, 16, 0, 1), (0, 4, 16, 0, 0), (0, 4, 16, 1, 1), (0, 4, 8, 0, 1), (0, 4, 32, 0, 1), (0, 4, 0, 0, 1), (0, 2, 16, 0, 1), (1, 4, 16, 0, 1)]
    for pair, wgs, sdiv, rej, pdl in configs:
        tune(4, pair); tune(0, wgs); tune(3, wgs); tune(1, sdiv); tune(2, rej); tune(5, pdl)
        out = ops.tc_score_topk(hb, wb, None, k, target=tgt)
        if not rej:
            if ref is None:
                ref = out
            else:
                assert torch.equal(ref["topk_idx"], out["topk_idx"]) and torch.equal(ref["topk_val"], out["topk_val"]), (pair, wgs, sdiv)
                assert torch.equal(ref["target_score"], out["target_score"])
        ms = timeit(lambda: ops.tc_score_topk(hb, wb, None, k, target=tgt))
        ms_probe = timeit(lambda: ops._lib.call("asme_b200_tc_score_pipeline_probe", ops._p(hb), R, hb.shape[1], ops._p(wb), V, 1, ops._stream()))
        print(json.dumps(dict(R=R, V=V, Kp=hb.shape[1], k=k, pair=pair, wgs=wgs, sample_div=sdiv, reject_all=rej, pdl=pdl, ms=round(ms, 4),
                              tflops=round(fl / ms / 1e9, 1), probe_ld_ms=round(ms_probe, 4))), flush=True)
    ts = ref["target_score"]
    refs = {}
    for pair, wgs in itertools.product((0,), (2, 4)):
        tune(4, pair); tune(0, wgs); tune(3, wgs); tune(1, 16); tune(2, 0); tune(5, 1)
        for name, fn in dict(count=lambda: ops.tc_score_topk(hb, wb, None, 0, target=tgt, target_score_in=ts, capture_target=False),
                             topk_count=lambda: ops.tc_score_topk(hb, wb, None, k, target=tgt, target_score_in=ts),
                             ce=lambda: ops.tc_score_ce_partial(hb, wb, None, tgt)).items():
            out = fn()
            if name == "count":
                key = (out["n_greater"], out["n_tie_lower"])
                if name in refs:
                    assert all(torch.equal(x, y) for x, y in zip(refs[name], key)), (name, pair, wgs)
                refs.setdefault(name, key)
            ms = timeit(fn)
            print(json.dumps(dict(R=R, V=V, Kp=hb.shape[1], case=name, pair=pair, wgs=wgs, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))), flush=True)
    tune(4, 0); tune(0, 4); tune(3, 4); tune(1, 16); tune(2, 0); tune(5, 1)


if __name__ == "__main__":
    shapes = [(1024, 1_000_003, 128, 10, False), (1024, 1_000_003, 128, 10, True), (4096, 1_000_003, 128, 10, False), (5253, 3709, 64, 10, False)]
    if len(sys.argv) > 1:
        shapes = shapes[:int(sys.argv[1])]
    for s in shapes:
        run(*s)
