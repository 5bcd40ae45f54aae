"""Vocab-sharded scoring on real ranks (``-m gpu``, needs >= 2 GPUs: ``gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py``;
skipped on a single-GPU box, where tests/test_gpu_tc.py covers the per-shard kernels and tests/test_sharded_cpu.py the exchange).

Every rank builds the same model, encodes ITS users and calls ``evaluate_rank_sharded`` (NCCL: one all-gather, one all-to-all); the
result must EQUAL the unsharded ``evaluate_rank`` of the same users on the same rank: top-k ids and scores, target scores and ranks
bit for bit (the lists are exact on both paths), the validation loss to fp32 rounding.  The sharded training cross entropy
(``sharded.sharded_ce``: all-reduce(MAX) + one packed all-reduce(SUM), dH all-reduced) is compared with the unsharded kernels."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, path):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from asme_b200 import ops, sharded
    from asme_b200.graphs import GraphedEvalStep
    from asme_b200.models import BERT4RecModel
    torch.manual_seed(0)
    V, S, H, B = 50021, 40, 128, 96
    model = BERT4RecModel(H, 2, 2, V, S, 0.0, initializer_range=0.05).cuda().eval()
    g = torch.Generator().manual_seed(100 + rank)
    seq = torch.randint(3, V, (B, S), generator=g)
    lengths = torch.randint(2, S, (B,), generator=g)
    seq = torch.where(torch.arange(S).unsqueeze(0) < lengths.unsqueeze(1), seq, torch.zeros_like(seq))
    seq[torch.arange(B), lengths] = 1
    seq = seq.cuda()
    free = torch.randint(3, V, (B,), generator=g).cuda()
    first = model.evaluate_rank(seq, seq.ne(0), {}, free, k=20)
    col = torch.randint(0, 20, (B,), generator=g).cuda()
    planted = first["topk_idx"].gather(1, col.unsqueeze(1)).squeeze(1).to(torch.int64)
    target = torch.where(torch.arange(B, device="cuda") % 2 == 0, planted, free)
    target[3] = 0                                                                     # an ignored row in the loss
    report = {}
    for full_rank in (False, True):
        want = model.evaluate_rank(seq, seq.ne(0), {}, target, k=10, with_loss=True, full_rank=full_rank)
        got = model.evaluate_rank_sharded(seq, seq.ne(0), {}, target, k=10, with_loss=True, full_rank=full_rank)
        tag = "full" if full_rank else "lite"
        report[f"{tag}.idx"] = bool(torch.equal(got["topk_idx"], want["topk_idx"]))
        report[f"{tag}.val"] = bool(torch.equal(got["topk_val"], want["topk_val"]))
        report[f"{tag}.ts"] = bool(torch.equal(got["target_score"], want["target_score"]))
        inside = want["rank"] <= 10
        report[f"{tag}.rank_inside"] = bool(torch.equal(got["rank"][inside], want["rank"][inside]))
        report[f"{tag}.rank_outside"] = bool((got["rank"][~inside] > 10).all()) and \
            float(((got["rank"][~inside] - want["rank"][~inside]).abs().double() / want["rank"][~inside].double()).max()) < (0.02 if full_rank else 1e-9)
        report[f"{tag}.loss"] = abs(float(got["loss"]) - float(want["loss"])) < 1e-5 * abs(float(want["loss"]))
        report[f"{tag}.n_inside"] = int(inside.sum())
    # the same step replayed from a CUDA graph with its NCCL exchanges captured
    graphed = GraphedEvalStep(lambda b: model.evaluate_rank_sharded(b["seq"], b["seq"].ne(0), {}, b["target"], k=10, with_loss=True))
    out = graphed({"seq": seq, "target": target})
    want = model.evaluate_rank(seq, seq.ne(0), {}, target, k=10, with_loss=True, full_rank=False)
    report["graph.idx"] = bool(torch.equal(out["topk_idx"], want["topk_idx"])) and bool(torch.equal(out["rank"], want["rank"]))

    # ---- vocab-sharded training cross entropy on the same rows everywhere
    gen = torch.Generator(device="cuda").manual_seed(7)               # same seed on every rank: the SAME rows
    R, Hh = 300, 64
    h = torch.randn(R, Hh, generator=gen, device="cuda")
    w = torch.randn(V, Hh, generator=gen, device="cuda") * 0.2
    b = torch.randn(V, generator=gen, device="cuda") * 0.1
    t = torch.randint(1, V, (R,), generator=gen, device="cuda")
    hb, wb = ops.cast_bf16(h), ops.cast_bf16(w)
    v0, v1 = sharded.shard_range(V, world, rank)
    ce = sharded.sharded_ce(lambda: ops.tc_score_ce_partial(hb, wb[v0:v1].contiguous(), b[v0:v1].clone(), t, v0=v0), t, pad_id=0,
                            rescale=ops.ce_rescale)
    rmax, rsum, tl = ops.tc_score_ce_partial(hb, wb, b, t)
    lse = rmax + torch.log(rsum)
    report["ce.lse"] = bool(torch.allclose(ce["lse"], lse, rtol=1e-6, atol=1e-6))
    report["ce.loss"] = abs(float(ce["loss"]) - float((lse - tl).mean())) < 1e-5
    dW_s, db_s = torch.zeros(v1 - v0, Hh, device="cuda"), torch.zeros(v1 - v0, device="cuda")
    dh = sharded.sharded_ce_backward(lambda l: ops.tc_score_ce_bwd(hb, wb[v0:v1].contiguous(), b[v0:v1].clone(), t, l, 1.0 / R, Hh, dW_s, db_s, v0=v0),
                                     ce["lse"])
    dW, db = torch.zeros(V, Hh, device="cuda"), torch.zeros(V, device="cuda")
    dh_ref = ops.tc_score_ce_bwd(hb, wb, b, t, lse, 1.0 / R, Hh, dW, db)
    report["ce.dh"] = float((dh - dh_ref).norm() / dh_ref.norm()) < 1e-5
    report["ce.dw"] = float((dW_s - dW[v0:v1]).norm() / dW[v0:v1].norm()) < 1e-5 and float((db_s - db[v0:v1]).norm() / db[v0:v1].norm()) < 1e-5
    torch.save(report, f"{path}.{rank}")
    torch.cuda.synchronize()
    os._exit(0)          # captured NCCL collectives keep the communicator busy: no orderly teardown


@pytest.mark.parametrize("world", [2])
def test_sharded_equals_unsharded_on_real_ranks(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    path = str(tmp_path / "res")
    ctx = mp.spawn(_worker, args=(world, _free_port(), path), nprocs=world, join=False)
    for p in ctx.processes:
        p.join(300)
    reports = [torch.load(f"{path}.{r}", weights_only=False) for r in range(world)]
    for r, rep in enumerate(reports):
        bad = [k for k, v in rep.items() if v is False]
        assert not bad, f"rank {r}: {bad} ({rep})"
        assert rep["lite.n_inside"] > 10
