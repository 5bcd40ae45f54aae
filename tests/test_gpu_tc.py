"""Parity tests of the tensor-core (tcgen05 / TMEM / TMA) scoring path against the CPU oracle.

Two kinds of inputs:
  * small-integer data: every product and partial sum is exactly representable, so the tensor-core result is
    bit-exact whatever its accumulation order, the scores contain MANY ties, and top-k ids / target scores /
    rank counts must equal the oracle's exactly (ties -> lowest item id);
  * gaussian data rounded to bf16: fp32-accumulation-order noise only; values within 1e-5 relative of a float64
    matmul of the same bf16 operands, ids exact wherever the oracle's k-th / (k+1)-th scores are separated.
Against the un-rounded fp32 operands the bf16 path is within 1e-3 relative of the logit scale (north_star's bf16
tolerance); that bound is asserted in test_tc_ce_matches_oracle.
"""
import numpy as np
import pytest
import torch

from oracle import asme_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from asme_b200 import ops
    return ops


def int_data(gen, R, V, H, lo=-3, hi=4):
    h = torch.randint(lo, hi, (R, H), generator=gen).float()
    w = torch.randint(lo, hi, (V, H), generator=gen).float()
    b = torch.randint(-2, 3, (V,), generator=gen).float()
    return h, w, b


def ref_logits(hb, wb, bias):
    """float64 scores of the bf16 operands (what the tensor cores see), as float32"""
    l = hb.float().double() @ wb.float().double().t()
    if bias is not None:
        l = l + bias.double()
    return l


@pytest.mark.parametrize("R,V,H,k", [(1, 13, 64, 5), (5, 255, 64, 10), (128, 256, 128, 10), (130, 257, 128, 1),
                                     (300, 3709, 64, 10), (77, 5000, 256, 32), (64, 1031, 100, 7), (256, 70001, 128, 10)])
@pytest.mark.parametrize("use_bias", [False, True])
def test_tc_topk_integer_exact(ops, R, V, H, k, use_bias):
    gen = torch.Generator().manual_seed(R * 7 + V + H + k)
    h, w, b = int_data(gen, R, V, H)
    bias = b if use_bias else None
    target = torch.randint(0, V, (R,), generator=gen)
    logits = ref_logits(h, w, bias).float().numpy()               # exact integers
    hb, wb = ops.cast_bf16(h.cuda()), ops.cast_bf16(w.cuda())
    assert hb.shape[1] == ops.padded_k(H) and torch.equal(hb[:, :H].float().cpu(), h)
    bias_d = None if bias is None else bias.cuda()
    out = ops.tc_score_topk(hb, wb, bias_d, k, target=target.cuda())
    kk = min(k, V)
    want_ids = O.topk_ids(logits, kk)
    got_ids = out["topk_idx"].cpu().numpy()
    got_val = out["topk_val"].cpu().numpy()
    np.testing.assert_array_equal(got_ids[:, :kk], want_ids)
    np.testing.assert_array_equal(got_val[:, :kk], np.take_along_axis(logits, want_ids, axis=1))
    if kk < k:
        assert (got_ids[:, kk:] == -1).all()
    st = logits[np.arange(R), target.numpy()]
    np.testing.assert_array_equal(out["target_score"].cpu().numpy(), st)
    # exact rank counts against the captured target score
    cnt = ops.tc_score_topk(hb, wb, bias_d, 0, target=target.cuda(), target_score_in=out["target_score"], capture_target=False)
    rank = (cnt["n_greater"] + cnt["n_tie_lower"] + 1).cpu().numpy()
    np.testing.assert_array_equal(rank, O.target_rank(logits, target.numpy()))
    # single sweep doing both
    both = ops.tc_score_topk(hb, wb, bias_d, k, target=target.cuda(), target_score_in=out["target_score"])
    np.testing.assert_array_equal(both["topk_idx"].cpu().numpy(), got_ids)
    np.testing.assert_array_equal((both["n_greater"] + both["n_tie_lower"] + 1).cpu().numpy(), rank)


def test_tc_topk_sharded_merge_exact(ops):
    """vocab-sharded scoring: per-shard top-k + owner-shard target score, merged (K-way) == unsharded oracle"""
    gen = torch.Generator().manual_seed(5)
    R, V, H, k, G = 200, 10007, 128, 10, 4
    h, w, b = int_data(gen, R, V, H)
    target = torch.randint(0, V, (R,), generator=gen)
    logits = ref_logits(h, w, b).float().numpy()
    hb = ops.cast_bf16(h.cuda())
    per = (V + G - 1) // G
    vals, idxs = [], []
    ts = torch.zeros(R, device="cuda")
    for g in range(G):
        v0, v1 = g * per, min(V, (g + 1) * per)
        wb = ops.cast_bf16(w[v0:v1].cuda())
        out = ops.tc_score_topk(hb, wb, b[v0:v1].clone().cuda(), k, target=target.cuda(), v0=v0)
        vals.append(out["topk_val"]); idxs.append(out["topk_idx"])
        ts += out["target_score"]                                    # all-reduce(SUM) in the multi-GPU path
    mv, mi = ops.topk_merge(torch.stack(vals), torch.stack(idxs), k)
    np.testing.assert_array_equal(mi.cpu().numpy(), O.topk_ids(logits, k))
    np.testing.assert_array_equal(ts.cpu().numpy(), logits[np.arange(R), target.numpy()])
    ng = torch.zeros(R, dtype=torch.int32, device="cuda")
    nt = torch.zeros_like(ng)
    for g in range(G):
        v0, v1 = g * per, min(V, (g + 1) * per)
        wb = ops.cast_bf16(w[v0:v1].cuda())
        c = ops.tc_score_topk(hb, wb, b[v0:v1].clone().cuda(), 0, target=target.cuda(), target_score_in=ts, v0=v0, capture_target=False)
        ng += c["n_greater"]; nt += c["n_tie_lower"]
    np.testing.assert_array_equal((ng + nt + 1).cpu().numpy(), O.target_rank(logits, target.numpy()))


def test_tc_topk_sampled_two_launch_path_integer_exact(ops):
    """long sweeps take the sample-then-threshold route (two launches); ids must still be exact, ties included"""
    gen = torch.Generator(device="cuda").manual_seed(3)
    R, V, H, k = 1000, 300_007, 64, 10
    h = torch.randint(-3, 4, (R, H), generator=gen, device="cuda").float()
    w = torch.randint(-3, 4, (V, H), generator=gen, device="cuda").float()
    b = torch.randint(-2, 3, (V,), generator=gen, device="cuda").float()
    target = torch.randint(0, V, (R,), generator=gen, device="cuda")
    hb, wb = ops.cast_bf16(h), ops.cast_bf16(w)
    out = ops.tc_score_topk(hb, wb, b, k, target=target)
    logits = h.double() @ w.double().t() + b.double()                       # exact small integers
    key = logits * float(1 << 20) - torch.arange(V, device="cuda", dtype=torch.float64)   # (score desc, id asc) as one exact key
    want = torch.topk(key, k, dim=1).indices
    assert torch.equal(out["topk_idx"].long(), want)
    assert torch.equal(out["topk_val"].double(), torch.gather(logits, 1, want))
    st = logits[torch.arange(R, device="cuda"), target]
    assert torch.equal(out["target_score"].double(), st)
    both = ops.tc_score_topk(hb, wb, b, k, target=target, target_score_in=out["target_score"])
    rank = (logits > st[:, None]).sum(1) + ((logits == st[:, None]) & (torch.arange(V, device="cuda")[None, :] < target[:, None])).sum(1) + 1
    assert torch.equal((both["n_greater"] + both["n_tie_lower"] + 1).long(), rank)
    assert torch.equal(both["topk_idx"], out["topk_idx"])


@pytest.mark.parametrize("R,V,H,k", [(100, 3709, 64, 10), (257, 20011, 128, 10)])
def test_tc_topk_gaussian(ops, R, V, H, k):
    gen = torch.Generator().manual_seed(V)
    h = torch.randn(R, H, generator=gen)
    w = torch.randn(V, H, generator=gen) * 0.1
    b = torch.randn(V, generator=gen) * 0.1
    target = torch.randint(0, V, (R,), generator=gen)
    hb, wb = ops.cast_bf16(h.cuda()), ops.cast_bf16(w.cuda())
    logits = ref_logits(hb[:, :H].cpu(), wb[:, :H].cpu(), b).numpy()          # float64
    out = ops.tc_score_topk(hb, wb, b.cuda(), k, target=target.cuda())
    got_ids = out["topk_idx"].cpu().numpy().astype(np.int64)
    got_val = out["topk_val"].cpu().numpy()
    order = np.argsort(-logits, axis=1, kind="stable")
    want_val = np.take_along_axis(logits, order[:, :k + 1], axis=1)
    scale = np.abs(logits).max()
    np.testing.assert_allclose(got_val, want_val[:, :k], rtol=1e-5, atol=1e-5 * scale)
    np.testing.assert_allclose(got_val, np.take_along_axis(logits, got_ids, axis=1), rtol=1e-5, atol=1e-5 * scale)
    gaps = want_val[:, :-1] - want_val[:, 1:]
    separated = gaps.min(axis=1) > 4e-6 * scale          # rows whose top-(k+1) scores are all distinct beyond fp32 noise
    assert separated.mean() > 0.5
    np.testing.assert_array_equal(got_ids[separated], order[separated, :k])
    np.testing.assert_allclose(out["target_score"].cpu().numpy(), logits[np.arange(R), target.numpy()], rtol=1e-5, atol=1e-5 * scale)


@pytest.mark.parametrize("R,V,H", [(1, 13, 64), (300, 3709, 64), (130, 12104, 64), (100, 1031, 100), (64, 30000, 128)])
def test_tc_ce_matches_oracle(ops, R, V, H):
    gen = torch.Generator().manual_seed(R + V)
    h = torch.randn(R, H, generator=gen)
    w = torch.randn(V, H, generator=gen) * 0.2
    b = torch.randn(V, generator=gen) * 0.1
    target = torch.randint(0, V, (R,), generator=gen)
    hb, wb = ops.cast_bf16(h.cuda()), ops.cast_bf16(w.cuda())
    rmax, rsum, tl = ops.tc_score_ce_partial(hb, wb, b.cuda(), target.cuda())
    nll = (rmax + torch.log(rsum) - tl).cpu().double()
    logits = ref_logits(hb[:, :H].cpu(), wb[:, :H].cpu(), b)
    want = torch.logsumexp(logits, dim=1) - logits[torch.arange(R), target]
    # same bf16 operands, float64 reference: fp32 accumulation noise only
    torch.testing.assert_close(nll, want, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(tl.cpu().double(), logits[torch.arange(R), target], rtol=1e-5, atol=2e-5)
    # un-rounded fp32 operands (the reference's arithmetic): bf16 tolerance of north_star, 1e-3 of the logit scale
    full = O.project(h, w, b).double()
    want32 = torch.logsumexp(full, dim=1) - full[torch.arange(R), target]
    assert (nll - want32).abs().max() <= 1e-3 * max(1.0, float(full.abs().max())) * 4
    assert abs(float(nll.mean() - want32.mean())) <= 1e-3 * float(want32.mean())


def test_tc_ce_sharded(ops):
    gen = torch.Generator().manual_seed(11)
    R, V, H, G = 150, 5003, 64, 3
    h = torch.randn(R, H, generator=gen)
    w = torch.randn(V, H, generator=gen) * 0.2
    b = torch.randn(V, generator=gen) * 0.1
    target = torch.randint(0, V, (R,), generator=gen)
    hb = ops.cast_bf16(h.cuda())
    per = (V + G - 1) // G
    parts = []
    for g in range(G):
        v0, v1 = g * per, min(V, (g + 1) * per)
        parts.append(ops.tc_score_ce_partial(hb, ops.cast_bf16(w[v0:v1].cuda()), b[v0:v1].clone().cuda(), target.cuda(), v0=v0))
    m = torch.stack([p[0] for p in parts]).max(dim=0).values               # all-reduce(MAX)
    s = sum(p[1] * torch.exp(p[0] - m) for p in parts)                     # rescale + all-reduce(SUM)
    tl = sum(p[2] for p in parts)
    nll = (m + torch.log(s) - tl).cpu().double()
    logits = ref_logits(hb[:, :H].cpu(), ops.cast_bf16(w.cuda())[:, :H].cpu(), b)
    want = torch.logsumexp(logits, dim=1) - logits[torch.arange(R), target]
    torch.testing.assert_close(nll, want, rtol=1e-5, atol=2e-5)


def test_tc_topk_c5_shape_properties(ops):
    """BASELINE config 5 at full size (1M-item catalog, hidden 128, 1024 users): size-independent properties plus a
    direct comparison on a row subset."""
    gen = torch.Generator(device="cuda").manual_seed(7)
    R, V, H, k = 1024, 1_000_003, 128, 10
    h = torch.randn(R, H, generator=gen, device="cuda")
    w = torch.randn(V, H, generator=gen, device="cuda") * 0.02
    b = (torch.rand(V, generator=gen, device="cuda") - 0.5) * 2e-3
    target = torch.randint(3, V, (R,), generator=gen, device="cuda")
    hb, wb = ops.cast_bf16(h), ops.cast_bf16(w)
    out = ops.tc_score_topk(hb, wb, b, k, target=target)
    val, idx, ts = out["topk_val"], out["topk_idx"].long(), out["target_score"]
    assert (val[:, :-1] >= val[:, 1:]).all()                                   # sortedness
    assert (idx >= 0).all() and (idx < V).all()
    assert all(len(set(r.tolist())) == k for r in idx[:64].cpu())              # distinct ids
    rec = (hb.float().unsqueeze(1) * wb[idx.reshape(-1)].float().view(R, k, -1)).sum(-1) + b[idx]
    torch.testing.assert_close(val, rec, rtol=1e-4, atol=1e-5)                 # every entry is a real score
    rec_t = (hb.float() * wb[target].float()).sum(-1) + b[target]
    torch.testing.assert_close(ts, rec_t, rtol=1e-4, atol=1e-5)
    # direct check on 32 rows against a dense fp32 matmul of the same bf16 operands
    sub = torch.arange(0, R, 32, device="cuda")
    dense = hb[sub].float() @ wb.float().t() + b
    dv, di = torch.topk(dense, k, dim=1)
    torch.testing.assert_close(val[sub], dv, rtol=1e-4, atol=1e-5)
    same = (di == idx[sub]).all(dim=1).float().mean()
    assert same > 0.9                                                           # fp32 accumulation-order near-ties aside
    # rank counts: consistent with the list (target in list <=> rank <= k) and with the dense scores
    cnt = ops.tc_score_topk(hb, wb, b, 0, target=target, target_score_in=ts, capture_target=False)
    rank = cnt["n_greater"] + cnt["n_tie_lower"] + 1
    in_list = (idx == target.unsqueeze(1)).any(dim=1)
    assert torch.equal(in_list, rank <= k)
    pos = (idx == target.unsqueeze(1)).float().argmax(dim=1) + 1
    assert torch.equal(pos[in_list].int(), rank[in_list])
    dense_rank = (dense > ts[sub].unsqueeze(1)).sum(dim=1) + 1
    assert ((dense_rank - rank[sub]).abs() <= 2).all()
    # predict path at full size: the log-sum-exp of the second sweep turns list entries into softmax scores (asme_b200.evaluation)
    rmax, rsum, _ = ops.tc_score_ce_partial(hb, wb, b, target)
    lse = rmax + torch.log(rsum)
    assert (lse >= val[:, 0]).all()                                            # lse >= max logit
    p = torch.exp(val - lse.unsqueeze(1))
    assert (p.sum(dim=1) <= 1.0 + 1e-5).all() and (p > 0).all()
    torch.testing.assert_close(lse[sub], torch.logsumexp(dense.double(), dim=1).float(), rtol=1e-5, atol=1e-4)
    # the candidate-FIFO depth (ring slots vs FIFO entries) must not change any result
    for depth in (8, 16):
        ops._lib.call("asme_b200_tc_score_tune", 7, depth)
        again = ops.tc_score_topk(hb, wb, b, k, target=target)
        assert torch.equal(again["topk_idx"], out["topk_idx"]) and torch.equal(again["topk_val"], val)
    ops._lib.call("asme_b200_tc_score_tune", 7, 12)


@pytest.mark.parametrize("R,V,H", [(1, 13, 64), (300, 3709, 64), (130, 12104, 64), (100, 1031, 100), (64, 30000, 128), (5253, 3709, 64),
                                   (200, 700, 256)])
def test_tc_ce_bwd_matches_autograd(ops, R, V, H):
    """dH, dW, dbias of the tensor-core CE backward vs torch autograd (float64) of mean cross-entropy over the same bf16 operands.
    dlogit passes through bf16 before the second MMA -> 2^-9 relative per element; compared in norm."""
    gen = torch.Generator(device="cuda").manual_seed(R + V + H)
    h = torch.randn(R, H, generator=gen, device="cuda")
    w = torch.randn(V, H, generator=gen, device="cuda") * 0.2
    b = torch.randn(V, generator=gen, device="cuda") * 0.1
    target = torch.randint(0, V, (R,), generator=gen, device="cuda")
    hb, wb = ops.cast_bf16(h), ops.cast_bf16(w)
    rmax, rsum, tl = ops.tc_score_ce_partial(hb, wb, b, target)
    lse = rmax + torch.log(rsum)
    dW = torch.zeros(V, H, device="cuda")
    db = torch.zeros(V, device="cuda")
    dh = ops.tc_score_ce_bwd(hb, wb, b, target, lse, 1.0 / R, H, dW, db)
    hd = hb[:, :H].double().requires_grad_(True)
    wd = wb[:, :H].double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    torch.nn.functional.cross_entropy(hd @ wd.t() + bd, target).backward()
    for name, got, want in (("dH", dh, hd.grad), ("dW", dW, wd.grad), ("dbias", db, bd.grad)):
        err = float((got.double() - want).norm() / want.norm())
        assert err < 6e-3, f"{name}: relative norm error {err:.5f}"
    # accumulation semantics: a second call adds to dW / dbias, overwrites dH
    dh2 = ops.tc_score_ce_bwd(hb, wb, b, target, lse, 1.0 / R, H, dW, db)
    assert torch.equal(dh, dh2)
    assert float((dW.double() - 2 * wd.grad).norm() / wd.grad.norm()) < 1.2e-2


def test_tc_ce_bwd_sharded(ops):
    gen = torch.Generator(device="cuda").manual_seed(21)
    R, V, H, G = 257, 5003, 64, 3
    h = torch.randn(R, H, generator=gen, device="cuda")
    w = torch.randn(V, H, generator=gen, device="cuda") * 0.2
    b = torch.randn(V, generator=gen, device="cuda") * 0.1
    target = torch.randint(0, V, (R,), generator=gen, device="cuda")
    hb, wb = ops.cast_bf16(h), ops.cast_bf16(w)
    rmax, rsum, _ = ops.tc_score_ce_partial(hb, wb, b, target)
    lse = rmax + torch.log(rsum)
    per = (V + G - 1) // G
    dh = torch.zeros(R, H, device="cuda")
    dW = torch.zeros(V, H, device="cuda")
    db = torch.zeros(V, device="cuda")
    for g in range(G):
        v0, v1 = g * per, min(V, (g + 1) * per)
        dWs, dbs = torch.zeros(v1 - v0, H, device="cuda"), torch.zeros(v1 - v0, device="cuda")
        dh += ops.tc_score_ce_bwd(hb, wb[v0:v1].contiguous(), b[v0:v1].clone(), target, lse, 1.0 / R, H, dWs, dbs, v0=v0)   # all-reduce(SUM)
        dW[v0:v1], db[v0:v1] = dWs, dbs
    hd = hb[:, :H].double().requires_grad_(True)
    wd = wb[:, :H].double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    torch.nn.functional.cross_entropy(hd @ wd.t() + bd, target).backward()
    for name, got, want in (("dH", dh, hd.grad), ("dW", dW, wd.grad), ("dbias", db, bd.grad)):
        assert float((got.double() - want).norm() / want.norm()) < 6e-3, name


@pytest.mark.parametrize("R,V,H,k", [(130, 5000, 128, 10), (64, 1031, 100, 7), (300, 3709, 64, 10)])
def test_tc_topk_bias_folded_into_the_contraction_exact(ops, R, V, H, k):
    """[h,1,1].[w,b_hi,b_lo] == h.w + b: with integer data the folded operands (Kp = H+2 rounded to 16, not a multiple of 64)
    must give exactly the ids / scores of the epilogue-bias path and of the oracle"""
    gen = torch.Generator().manual_seed(R + V + H)
    h, w, b = int_data(gen, R, V, H)
    b = b + 0.5                                                     # 0.5 steps: still exact in bf16 (hi part, lo = 0)
    target = torch.randint(0, V, (R,), generator=gen)
    logits = ref_logits(h, w, b).float().numpy()
    hb, wb = ops.cast_bf16_ext(h.cuda()), ops.cast_bf16_ext(w.cuda(), b.cuda())
    assert hb.shape[1] == wb.shape[1] == (H + 2 + 15) // 16 * 16
    out = ops.tc_score_topk(hb, wb, None, k, target=target.cuda())
    want_ids = O.topk_ids(logits, k)
    np.testing.assert_array_equal(out["topk_idx"].cpu().numpy(), want_ids)
    np.testing.assert_array_equal(out["topk_val"].cpu().numpy(), np.take_along_axis(logits, want_ids, axis=1))
    np.testing.assert_array_equal(out["target_score"].cpu().numpy(), logits[np.arange(R), target.numpy()])
    rmax, rsum, tl = ops.tc_score_ce_partial(hb, wb, None, target.cuda())
    want = torch.logsumexp(torch.from_numpy(logits).double(), dim=1)
    torch.testing.assert_close((rmax + torch.log(rsum)).cpu().double(), want, rtol=1e-5, atol=1e-4)
    # a real-valued bias: hi + lo keeps 16 mantissa bits
    bias = torch.randn(V, generator=gen) * 0.3
    wb2 = ops.cast_bf16_ext(w.cuda(), bias.cuda())
    got = ops.tc_score_topk(hb, wb2, None, k, target=target.cuda())["target_score"].cpu()
    want = torch.from_numpy(ref_logits(h, w, None).float().numpy())[torch.arange(R), target] + bias[target]
    torch.testing.assert_close(got, want, rtol=0, atol=2e-5)


@pytest.mark.parametrize("cap,live,V,H,plan", [(1000, 300, 3709, 64, 300), (1000, 1, 700, 64, 0), (640, 640, 1031, 128, 100), (2000, 129, 5003, 64, 2000)])
def test_tc_ce_with_device_row_count(ops, cap, live, V, H, plan):
    """row selections: the buffers have a capacity, the live count is in device memory.  Forward partials and the loss of the live
    rows equal the call on exactly those rows bit for bit (rows are independent in the sweep); the backward divides by the count on
    the device and matches autograd of the mean over the live rows; dead rows hold NaNs on purpose and must not leak."""
    gen = torch.Generator(device="cuda").manual_seed(cap + live)
    h = torch.randn(cap, H, generator=gen, device="cuda")
    h[live:] = float("nan")
    w = torch.randn(V, H, generator=gen, device="cuda") * 0.2
    b = torch.randn(V, generator=gen, device="cuda") * 0.1
    target = torch.randint(1, V, (cap,), generator=gen, device="cuda")
    target[live:] = 0
    n = torch.tensor([live], dtype=torch.int32, device="cuda")
    wb = ops.cast_bf16(w)
    hb = ops.cast_bf16(h, n_live=n)
    assert bool((hb[live:] == 0).all())
    rmax, rsum, tl = ops.tc_score_ce_partial(hb, wb, b, target, n_live=n, plan_rows=plan)
    hb_x = ops.cast_bf16(h[:live].contiguous())
    rmax_x, rsum_x, tl_x = ops.tc_score_ce_partial(hb_x, wb, b, target[:live].contiguous())
    # the split of the catalog over CTAs follows the planned row count, and with it the order in which the partial sums meet
    assert torch.equal(rmax[:live], rmax_x) and torch.equal(tl[:live], tl_x)
    torch.testing.assert_close(rsum[:live], rsum_x, rtol=2e-6, atol=0)
    acc = torch.zeros(2, device="cuda")
    lse = ops.ce_loss_from_partials(rmax, rsum, tl, acc[0:1], n, acc[1:2])
    acc_x = torch.zeros(1, device="cuda")
    lse_x = ops.ce_loss_from_partials(rmax_x, rsum_x, tl_x, acc_x)
    torch.testing.assert_close(lse[:live], lse_x, rtol=1e-6, atol=1e-6)
    assert float(acc[1]) == pytest.approx(float(acc_x[0]) / live, rel=1e-5)
    dW, db = torch.zeros(V, H, device="cuda"), torch.zeros(V, device="cuda")
    dh = ops.tc_score_ce_bwd(hb, wb, b, target, lse, 1.0, H, dW, db, n_live=n)
    hd = hb_x[:, :H].double().requires_grad_(True)
    wd = wb[:, :H].double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    torch.nn.functional.cross_entropy(hd @ wd.t() + bd, target[:live]).backward()
    for name, got, want in (("dH", dh[:live], hd.grad), ("dW", dW, wd.grad), ("dbias", db, bd.grad)):
        assert bool(torch.isfinite(got).all()), name
        err = float((got.double() - want).norm() / want.norm())
        assert err < 6e-3, f"{name}: relative norm error {err:.5f}"
