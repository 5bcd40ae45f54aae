"""One small invocation of the hot path on cuda:0, checked against the CPU oracle (driver's smoke())."""
import os
import sys

import numpy as np
import torch


def run(verbose: bool = True):
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import asme_oracle as O          # the checker, never the thing measured
    from asme_b200.models import BERT4RecModel
    from asme_b200.data import InputSequence

    torch.manual_seed(0)
    B, S, V, H, L, heads = 8, 24, 203, 64, 2, 2
    model = BERT4RecModel(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L, item_vocab_size=V,
                          max_seq_length=S, transformer_dropout=0.0, initializer_range=0.2)
    w_cpu = {k: v.detach().clone() for k, v in model.state_dict().items() if k != "_projection_layer.embedding.weight"}
    gen = torch.Generator().manual_seed(1)
    seq = torch.randint(3, V, (B, S), generator=gen)
    lengths = torch.randint(2, S, (B,), generator=gen)
    target = torch.zeros_like(seq)
    for i in range(B):
        seq[i, lengths[i]:] = 0
        m = torch.rand(int(lengths[i]), generator=gen) < 0.3
        m[0] = True
        target[i, :lengths[i]][m] = seq[i, :lengths[i]][m]
        seq[i, :lengths[i]][m] = 1
    model = model.cuda().train()
    model.precision = "fp32"      # strict-parity pass first (1e-4 / exact ranks up to fp32 near-ties), tensor-core pass below
    seq_d, target_d = seq.cuda(), target.cuda()

    # training step: fused CE loss + backward
    loss, ctx = model.loss_ce(seq_d, seq_d.ne(0), {}, target_d)
    model.loss_ce_backward(ctx)
    leaves = {k: v.clone().requires_grad_(True) for k, v in w_cpu.items()}
    ref_loss = O.cross_entropy_ignore_pad(O.bert4rec_logits(leaves, seq, heads, L), target)
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 1e-4 * max(1.0, abs(float(ref_loss))), (float(loss), float(ref_loss))
    g = dict(model.named_parameters())["_sequence_embedding_layer.item_embedding.embedding.weight"].grad.cpu()
    torch.testing.assert_close(g, leaves["_sequence_embedding_layer.item_embedding.embedding.weight"].grad, rtol=1e-3, atol=1e-5)

    # evaluation: one MASK per row -> fused scoring + top-k + rank vs the oracle's stable sort
    model.eval()
    ev = seq.clone()
    ev[ev == 1] = 5
    for i in range(B):
        ev[i, lengths[i]] = 1
    tgt = torch.randint(3, V, (B,), generator=gen)
    out = model.evaluate_rank(ev.cuda(), ev.cuda().ne(0), {}, tgt.cuda(), k=10)
    rows = O.select_masked_rows(O.bert4rec_logits(w_cpu, ev, heads, L), ev).numpy()
    want_rank = O.target_rank(rows, tgt.numpy())
    got_rank = out["rank"].cpu().numpy()
    # fp32 summation order differs between CPU and GPU: ranks may differ only where scores tie within 1e-5
    srt = np.sort(rows, axis=1)[:, ::-1]
    assert np.abs(np.take_along_axis(rows, out["topk_idx"].cpu().numpy().astype(np.int64), 1) - srt[:, :10]).max() < 1e-4
    assert (np.abs(got_rank - want_rank) <= 1).all() and (got_rank == want_rank).mean() >= 0.75, (got_rank, want_rank)
    logits = model(InputSequence(ev.cuda(), ev.cuda().ne(0), {}))
    assert logits.shape == (B, S, V)

    # the same two steps under the tensor-core policy (tcgen05 GEMMs + catalog scoring, bf16 operands): bf16 tolerances
    model.precision = "bf16"
    model.train()
    loss16, ctx16 = model.loss_ce(seq_d, seq_d.ne(0), {}, target_d)
    model.loss_ce_backward(ctx16)
    assert abs(float(loss16) - float(ref_loss)) < 1e-3 * max(1.0, abs(float(ref_loss))), (float(loss16), float(ref_loss))
    model.eval()
    out16 = model.evaluate_rank(ev.cuda(), ev.cuda().ne(0), {}, tgt.cuda(), k=10)
    # bf16 operands through two encoder layers with N(0, 0.2) weights: scores within 3e-2 of the logit scale; ranks may move
    # only past competitors whose oracle score is that close to the target's
    scale = np.abs(rows).max()
    tol = 3e-2 * scale
    got16 = np.take_along_axis(rows, out16["topk_idx"].cpu().numpy().astype(np.int64), 1)
    assert np.abs(got16 - srt[:, :10]).max() < tol, (np.abs(got16 - srt[:, :10]).max(), scale)
    st = rows[np.arange(B), tgt.numpy()][:, None]
    near = (np.abs(rows - st) <= tol).sum(axis=1) - 1
    assert (np.abs(out16["rank"].cpu().numpy() - want_rank) <= near).all(), (out16["rank"].cpu().numpy(), want_rank, near)
    if verbose:
        print(f"[asme_b200 smoke] fp32 loss={float(loss):.6f} bf16/tcgen05 loss={float(loss16):.6f} (oracle {float(ref_loss):.6f}); "
              f"ranks={got_rank.tolist()} OK")


if __name__ == "__main__":
    run()
