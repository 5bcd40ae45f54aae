"""cost of the exact top-k on the C5 scoring shape: candidate sweep with 10 / 20 / 32 entries, re-score + certificate, flagged pass
    python tools/perf_exact.py [R]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200"))
from asme_b200 import models, ops  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
V, H = (int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_003), 128
g = torch.Generator(device="cuda").manual_seed(0)
h = torch.randn(R, H, generator=g, device="cuda")
w = torch.randn(V, H, generator=g, device="cuda") * 0.02
b = torch.randn(V, generator=g, device="cuda") * 0.001
target = torch.randint(0, V, (R,), generator=g, device="cuda")
wb = ops.cast_bf16_ext(w, b)
hb = ops.cast_bf16_ext(h)
nb = ops.table_norm_bound(w, b)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for kc in (10, 20, 32):
    print(f"sweep, {kc}-entry lists: {timed(lambda: ops.tc_score_topk(hb, wb, None, kc, target=target)):.4f} ms")
print(f"candidate sweep (k=10 lists, union of the splits -> 64): {timed(lambda: ops.tc_score_candidates(hb, wb, None, 10, 64)):.4f} ms")
ops._lib.call("asme_b200_tc_score_tune", 8, 0)
print(f"candidate sweep, knob 8 = 0 (16 splits per row tile) -> 64: {timed(lambda: ops.tc_score_candidates(hb, wb, None, 10, 64)):.4f} ms")
ops._lib.call("asme_b200_tc_score_tune", 8, 1)
print(f"candidate sweep (k=10 lists, union of the splits -> 32): {timed(lambda: ops.tc_score_candidates(hb, wb, None, 10, 32)):.4f} ms")
o32 = ops.tc_score_candidates(hb, wb, None, 10, 32)
r32 = ops.topk_rescore(h, w, b, o32["cand_idx"], o32["cand_val"], 10, nb, target, cand_bound=o32["bound"])
print(f"re-score + certificate (32 candidates): {timed(lambda: ops.topk_rescore(h, w, b, o32['cand_idx'], o32['cand_val'], 10, nb, target, cand_bound=o32['bound'])):.4f} ms, uncertified rows {int(r32['n_flagged'])}")
o = ops.tc_score_candidates(hb, wb, None, 10, 64)
o = dict(topk_idx=o["cand_idx"], topk_val=o["cand_val"], bound=o["bound"])
print(f"re-score + certificate (64 candidates, k=10): {timed(lambda: ops.topk_rescore(h, w, b, o['topk_idx'], o['topk_val'], 10, nb, target, cand_bound=o['bound'])):.4f} ms")
r = ops.topk_rescore(h, w, b, o["topk_idx"], o["topk_val"], 10, nb, target, cand_bound=o["bound"])
print("uncertified rows:", int(r["n_flagged"]))
print(f"flagged pass (nothing flagged): {timed(lambda: ops.score_topk_flagged(h, w, b, target, r['target_score'], 10, r['row_flag'], r['topk_val'], r['topk_idx'], r['rank'])):.4f} ms")
print(f"whole exact call: {timed(lambda: models.score_rows_tc_exact(h, hb, wb, w, b, nb, target, 10)):.4f} ms")
print(f"bf16-only call:   {timed(lambda: models.score_rows_tc(hb, wb, None, target, 10, False)):.4f} ms")
print(f"table norm bound: {timed(lambda: ops.table_norm_bound(w, b)):.4f} ms")
flag = r["row_flag"].clone(); flag[::97] = 1
print(f"flagged pass ({int(flag.sum())} rows flagged): {timed(lambda: ops.score_topk_flagged(h, w, b, target, r['target_score'], 10, flag, r['topk_val'], r['topk_idx'], r['rank']), 3):.4f} ms")

# ---- bias kept out of the contraction: plain (V, H) table + per-chunk bias bounds (K = 128 instead of 144) -------------------------
wp = ops.cast_bf16(w)
hp = ops.cast_bf16(h, ld_out=H)
bbs = ops.bias_chunk_bounds(b)
print(f"candidate sweep, plain table + bias bounds: {timed(lambda: ops.tc_score_candidates(hp, wp, b, 10, 64, bias_bounds=bbs)):.4f} ms")
print(f"candidate sweep, plain table, bias added in every chunk: {timed(lambda: ops.tc_score_candidates(hp, wp, b, 10, 64)):.4f} ms")
print(f"candidate sweep, plain table, no bias at all: {timed(lambda: ops.tc_score_candidates(hp, wp, None, 10, 64)):.4f} ms")
o2 = ops.tc_score_candidates(hp, wp, b, 10, 64, bias_bounds=bbs)
r2 = ops.topk_rescore(h, w, b, o2["cand_idx"], o2["cand_val"], 10, nb, target, cand_bound=o2["bound"])
print("uncertified rows (bounds path):", int(r2["n_flagged"]), "| lists equal the folded path's:", bool(torch.equal(r2["topk_idx"], r["topk_idx"])),
      bool(torch.equal(r2["topk_val"], r["topk_val"])))
print(f"whole exact call, bounds path: {timed(lambda: models.score_rows_tc_exact(h, hp, wp, w, b, nb, target, 10, bias_bounds=(b, bbs))):.4f} ms")
bl = torch.randn(V, generator=g, device="cuda") * 0.5          # a bias as large as the scores' spread (popularity-like)
bbl = ops.bias_chunk_bounds(bl)
nbl = ops.table_norm_bound(w, bl)
print(f"candidate sweep, bounds path, LARGE bias (sigma 0.5): {timed(lambda: ops.tc_score_candidates(hp, wp, bl, 10, 64, bias_bounds=bbl)):.4f} ms")
o3 = ops.tc_score_candidates(hp, wp, bl, 10, 64, bias_bounds=bbl)
r3 = ops.topk_rescore(h, w, bl, o3["cand_idx"], o3["cand_val"], 10, nbl, target, cand_bound=o3["bound"])
dense = (h[:64].double() @ w.double().t() + bl.double())
want = torch.sort(dense, dim=1, descending=True, stable=True).indices[:, :10]
print("large bias: uncertified rows", int(r3["n_flagged"]), "| top-10 of the first 64 rows equals dense fp64:", bool(torch.equal(want.int(), r3["topk_idx"][:64])))
