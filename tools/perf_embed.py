"""Embedding gather-and-sum micro-benchmark (CUDA events, warm; rotating id sets so the ids are not L2-resident).
Algorithmic bytes per token: n_tables*H*4 read + n_ids*8 read + H*4 written (positions are generated)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")]
import torch
from asme_b200 import ops
from tools.perf_score import timeit

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def run(B, S, V, H, pos, ln, p, nxt=False):
    g = torch.Generator(device="cuda").manual_seed(0)
    table = torch.randn(V, H, device="cuda", generator=g)
    P = torch.randn(S, H, device="cuda", generator=g) if pos else None
    gam, bet = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
    ids = [torch.randint(0, V, (B, S), device="cuda", generator=g) for _ in range(4)]
    specs = [ops.EmbedSpec(i, table, pos_table=P, ln1=(gam, bet) if ln else None, p_drop=p, seed=7) for i in ids]
    it = [0]

    def step():
        ops.embed_fwd(specs[it[0] % 4], B, S, next_ln=(gam, bet) if nxt else None)
        it[0] += 1
    ms = timeit(step, iters=20)
    T = B * S
    nbytes = T * ((1 + int(pos)) * H * 4 + 8 + H * 4 + (H * 2 if nxt else 0))
    print(json.dumps(dict(T=T, V=V, H=H, pos=pos, ln=ln, p=p, next_ln=nxt, ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1),
                          frac=round(nbytes / ms / 1e6 / PEAK, 3))), flush=True)


def run_ln(M, H):
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.randn(M, H, device="cuda", generator=g) for _ in range(3)]
    gam, bet = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
    it = [0]

    def step():
        ops.layernorm_fwd_bf16(xs[it[0] % 3], gam, bet)
        it[0] += 1
    ms = timeit(step, iters=20)
    nbytes = M * H * 6
    print(json.dumps(dict(kernel="layernorm_fwd_bf16", M=M, H=H, ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1),
                          frac=round(nbytes / ms / 1e6 / PEAK, 3))), flush=True)


if __name__ == "__main__":
    wide = int(os.environ.get("ROW_WIDE", "1"))
    ops._lib.call("asme_b200_rowwise_tune", 0, wide)
    print(json.dumps(dict(wide_rows=wide)), flush=True)
    run(1024, 200, 1_000_003, 128, False, True, 0.0, nxt=True)      # the C5 evaluation step's call (BERT4Rec: no positions)
    run(256, 200, 3709, 64, False, True, 0.2, nxt=True)             # the C2 training step's call
    run_ln(204800, 128)
    run_ln(51200, 64)
    for (B, S, V, H) in [(1024, 200, 1_000_003, 128), (4096, 200, 1_000_003, 128), (256, 200, 3709, 64), (4096, 50, 13047, 64)]:
        for pos, ln, p in [(False, False, 0.0), (False, True, 0.0), (True, True, 0.0), (False, True, 0.2)]:
            run(B, S, V, H, pos, ln, p)
        run(B, S, V, H, True, True, 0.0, nxt=True)
