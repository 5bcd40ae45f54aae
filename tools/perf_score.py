"""Micro-benchmark of the tensor-core scoring kernels (CUDA events, warm, HBM-resident operands).
    python tools/perf_score.py [R V H k] ...
"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")]
import torch
from asme_b200 import ops


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(R, V, H, k):
    gen = torch.Generator(device="cuda").manual_seed(0)
    h = torch.randn(R, H, device="cuda", generator=gen)
    w = torch.randn(V, H, device="cuda", generator=gen) * 0.02
    b = torch.zeros(V, device="cuda")
    tgt = torch.randint(3, V, (R,), device="cuda", generator=gen)
    hb, wb = ops.cast_bf16(h), ops.cast_bf16(w)
    fl = 2.0 * R * V * hb.shape[1]
    res = {}
    out = ops.tc_score_topk(hb, wb, b, k, target=tgt)
    ts = out["target_score"]
    cases = {
        "topk+bias": lambda: ops.tc_score_topk(hb, wb, b, k, target=tgt),
        "topk": lambda: ops.tc_score_topk(hb, wb, None, k, target=tgt),
        "count": lambda: ops.tc_score_topk(hb, wb, None, 0, target=tgt, target_score_in=ts, capture_target=False),
        "topk+count": lambda: ops.tc_score_topk(hb, wb, None, k, target=tgt, target_score_in=ts),
        "ce": lambda: ops.tc_score_ce_partial(hb, wb, None, tgt),
        "topk+bias_folded": (lambda hbf=ops.cast_bf16_ext(h), wbf=ops.cast_bf16_ext(w, b): ops.tc_score_topk(hbf, wbf, None, k, target=tgt)),
        "probe": lambda: ops._lib.call("asme_b200_tc_score_pipeline_probe", ops._p(hb), R, hb.shape[1], ops._p(wb), V, 0, ops._stream()),
        "probe_ld": lambda: ops._lib.call("asme_b200_tc_score_pipeline_probe", ops._p(hb), R, hb.shape[1], ops._p(wb), V, 1, ops._stream()),
        "cast_w": lambda: ops.cast_bf16(w),
    }
    only = os.environ.get("CASES")
    for name, fn in cases.items():
        if only and name not in only.split(","):
            continue
        ms = timeit(fn)
        res[name] = dict(ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1) if name != "cast_w" else None)
    print(json.dumps(dict(R=R, V=V, H=H, k=k, **res)))


if __name__ == "__main__":
    args = [int(x) for x in sys.argv[1:]]
    shapes = [tuple(args[i:i + 4]) for i in range(0, len(args), 4)] or [(1024, 1_000_003, 128, 10), (4096, 1_000_003, 128, 10),
                                                                         (1024, 1_000_003, 64, 10), (5253, 3709, 64, 10), (256, 3709, 64, 10)]
    for s in shapes:
        run(*s)
