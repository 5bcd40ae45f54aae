"""Parity of the tensor-core attention kernels against the fp32 SIMT kernels (themselves pinned to the oracle's
restatement of Attention.forward, transformer_layers.py:145-155, in test_gpu_kernels.py) on the same bf16-rounded q/k/v,
with identical dropout masks.  Tolerance: the probabilities are rounded to bf16 before the second MMA and the
context is stored as bf16, i.e. 2^-8 relative per element -> 1e-2 of the output scale element-wise, 3e-3 in norm."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [  # B, S, heads, d
    (3, 200, 2, 32), (2, 50, 2, 32), (2, 256, 1, 64), (2, 37, 4, 16), (1, 130, 4, 32), (5, 1, 2, 32), (2, 129, 2, 64)]


@pytest.fixture(scope="module", params=["default", "forward_v2_everywhere"])
def ops(request):
    """every test of this file runs twice: with the launcher's own choice of the forward kernel (first kernel for grids smaller than
    the machine, as all of these are) and with the second kernel -- probabilities in tensor memory as the A operand of the second
    MMA, two CTAs per SM -- forced for every shape (knob 2 = 3)"""
    from asme_b200 import ops
    ops._lib.call("asme_b200_tc_attn_tune", 2, 3 if request.param == "forward_v2_everywhere" else 2)
    yield ops
    ops._lib.call("asme_b200_tc_attn_tune", 2, 2)


def make(gen, B, S, heads, d, pad="right"):
    H = heads * d
    qkv = (torch.randn(B * S, 3 * H, generator=gen, device="cuda") * 0.7).bfloat16()
    lengths = torch.randint(1, S + 1, (B,), generator=gen, device="cuda")
    lengths[0] = S
    pos = torch.arange(S, device="cuda").unsqueeze(0)
    valid = pos < lengths.unsqueeze(1) if pad == "right" else pos >= (S - lengths).unsqueeze(1)
    return qkv, valid


def unpack_keep(keep, B, heads, S):
    """(B*heads*S, 8) int32 keep bits -> (B, heads, S, S) bool"""
    bits = (keep.view(B, heads, S, 8, 1) >> torch.arange(32, device=keep.device, dtype=torch.int32)) & 1
    return bits.reshape(B, heads, S, 256)[..., :S].bool()


def torch_attention(qkv, valid, B, S, heads, causal, keep=None, p_drop=0.0):
    """fp32 restatement of Attention.forward (transformer_layers.py:145-155) with an explicit dropout keep mask"""
    H = qkv.shape[1] // 3
    d = H // heads
    x = qkv.float().view(B, S, 3, heads, d).permute(2, 0, 3, 1, 4)          # (3,B,heads,S,d)
    q, k, v = x[0], x[1], x[2]
    scores = q @ k.transpose(-1, -2) / d ** 0.5
    mask = torch.ones(B, 1, S, S, dtype=torch.bool, device=qkv.device)
    if valid is not None:
        mask = mask & valid.view(B, 1, 1, S)
    if causal:
        mask = mask & torch.tril(torch.ones(S, S, dtype=torch.bool, device=qkv.device))
    p = torch.softmax(scores.masked_fill(~mask, -1e9), dim=-1)
    if keep is not None:
        p = p * keep.float() / (1.0 - p_drop)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * S, H)


def norm_err(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("pad", ["right", "left"])
def test_tc_attn_fwd_scores_far_above_the_first_block(ops, causal, pad):
    """Rows whose largest score sits far (~ +90) above everything in the first key blocks, late in the sequence: the softmax of the
    tensor-core forward (exp2 of score - row maximum, bf16 probabilities in tensor memory) and its saved statistics must be those of
    the reference softmax (transformer_layers.py:145-155), next to ordinary sequences in the same launch, with right and with left
    padding (first blocks of padding keys only), causal or not."""
    B, S, heads, d = 4, 200, 2, 32
    H = heads * d
    gen = torch.Generator(device="cuda").manual_seed(11)
    qkv, valid = make(gen, B, S, heads, d, pad=pad)
    x = qkv.float().view(B, S, 3, heads, d)
    for b, q_rows, key in ((1, slice(60, 140), 185), (2, slice(0, 200), 199), (2, slice(150, 200), 97)):
        valid[b] = True
        u = torch.randn(d, generator=gen, device="cuda")
        u = u / u.norm()
        x[b, q_rows, 0, 1] += 22.0 * u                          # queries of head 1 ...
        x[b, key, 1, 1] = 23.0 * u                              # ... and one key: score ~ 22 * 23 / sqrt(32) = 89
    qkv = x.view(B * S, 3 * H).bfloat16()
    _, rst = ops.attn_fwd(qkv.float(), valid, B, S, heads, causal, 0.0, 77, 19, save_stats=True)
    ref = torch_attention(qkv, valid, B, S, heads, causal)
    got, gst, _ = ops.tc_attn_fwd(qkv, valid, B, S, heads, causal, 0.0, 77, 19, save_stats=True)
    assert norm_err(got, ref) < 4e-3
    torch.testing.assert_close(got.float(), ref, rtol=2e-2, atol=2e-2 * float(ref.abs().max()))
    torch.testing.assert_close(gst[0], rst[0], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gst[1], rst[1], rtol=1e-4, atol=1e-5)
    again, gst2, _ = ops.tc_attn_fwd(qkv, valid, B, S, heads, causal, 0.0, 77, 19, save_stats=True)
    assert torch.equal(got, again) and torch.equal(gst, gst2)


@pytest.mark.parametrize("p_drop", [0.5, 0.25, 0.2, 0.1, 0.01])
def test_tc_attn_dropout_stream_is_bernoulli(ops, p_drop):
    """The keep bits are drawn bit-sliced (one hash word per bit plane of 32 keys' uniform numbers, attention_tc.cu): check the
    keep rate (5 sigma), the independence of adjacent keys, of keys 32 apart (same bit of adjacent words), of adjacent
    queries and of adjacent heads, and that two sites / seeds are unrelated.  Reference: nn.Dropout on the attention
    probabilities (transformer_layers.py:152-153) -- Bernoulli(1 - p) per element, kept values scaled by 1 / (1 - p)."""
    B, S, heads, d = 8, 256, 4, 16
    gen = torch.Generator(device="cuda").manual_seed(3)
    qkv, _ = make(gen, B, S, heads, d)
    _, _, keep = ops.tc_attn_fwd(qkv, None, B, S, heads, False, p_drop, 1234, 5, save_stats=True)
    km = unpack_keep(keep, B, heads, S).double()                   # (B, heads, S, S)
    n = km.numel()
    q = 1.0 - round(p_drop * 65536) / 65536
    var = q * (1 - q)
    assert abs(float(km.mean()) - q) < 5 * (var / n) ** 0.5
    c = km - q
    tol = 5 / n ** 0.5
    assert abs(float((c[..., 1:] * c[..., :-1]).mean()) / var) < tol            # adjacent keys
    assert abs(float((c[..., 32:] * c[..., :-32]).mean()) / var) < tol          # same bit, adjacent 32-key words
    assert abs(float((c[:, :, 1:] * c[:, :, :-1]).mean()) / var) < tol          # adjacent queries
    assert abs(float((c[:, 1:] * c[:, :-1]).mean()) / var) < tol                # adjacent heads
    assert float((km.mean(dim=(0, 1, 2)) - q).abs().max()) < 6 * (var / (n / S)) ** 0.5   # no key position is special
    for seed, site in ((1235, 5), (1234, 6)):
        _, _, other = ops.tc_attn_fwd(qkv, None, B, S, heads, False, p_drop, seed, site, save_stats=True)
        co = unpack_keep(other, B, heads, S).double() - q
        assert abs(float((c * co).mean()) / var) < tol


@pytest.mark.parametrize("B,S,heads,d", CASES)
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("p_drop", [0.0, 0.2])
def test_tc_attn_fwd_matches_simt(ops, B, S, heads, d, causal, p_drop):
    gen = torch.Generator(device="cuda").manual_seed(B * 1000 + S + heads + d)
    qkv, valid = make(gen, B, S, heads, d)
    _, rst = ops.attn_fwd(qkv.float(), valid, B, S, heads, causal, 0.0, 77, 19, save_stats=True)
    got, gst, keep = ops.tc_attn_fwd(qkv, valid, B, S, heads, causal, p_drop, 77, 19, save_stats=True)
    if p_drop == 0.0:
        ref, _ = ops.attn_fwd(qkv.float(), valid, B, S, heads, causal)                # the fp32 SIMT kernel
        assert keep is None
    else:
        # the tensor-core kernel has its own dropout stream and hands the keep bits to the backward pass: compare with the
        # reference arithmetic under exactly that mask, and check the mask is Bernoulli(1 - p)
        km = unpack_keep(keep, B, heads, S)
        ref = torch_attention(qkv, valid, B, S, heads, causal, km, p_drop)
        if B * heads * S * S >= 20000:
            # ... on the keys that can be attended: chunks of padding keys only carry no dropout stream (their probabilities
            # are exactly zero, the kernel skips them and leaves their keep bits zero)
            attendable = valid.view(B, 1, 1, S).expand(B, heads, S, S)
            assert abs(float(km[attendable].float().mean()) - (1.0 - p_drop)) < 0.02
        again, _, keep2 = ops.tc_attn_fwd(qkv, valid, B, S, heads, causal, p_drop, 77, 19, save_stats=True)
        assert torch.equal(keep, keep2) and torch.equal(got, again)                 # pure function of (seed, site, element)
        other, _, keep3 = ops.tc_attn_fwd(qkv, valid, B, S, heads, causal, p_drop, 78, 19, save_stats=True)
        assert not torch.equal(keep, keep3)
    assert norm_err(got, ref) < 4e-3
    torch.testing.assert_close(got.float(), ref, rtol=2e-2, atol=2e-2 * float(ref.abs().max()))
    torch.testing.assert_close(gst[0], rst[0], rtol=1e-5, atol=1e-5)          # row max: fp32 accumulation noise only
    torch.testing.assert_close(gst[1], rst[1], rtol=1e-4, atol=1e-5)          # row sum-exp


def test_tc_attn_fwd_no_mask_and_fully_masked_rows(ops):
    """bidirectional without a padding mask (key_valid = NULL) and left padding under a causal mask: rows left of the first
    real token see no valid key and must attend uniformly to all S keys (-1e9 fill, quirk Q4)"""
    gen = torch.Generator(device="cuda").manual_seed(4)
    B, S, heads, d = 3, 70, 2, 32
    qkv, valid = make(gen, B, S, heads, d, pad="left")
    ref, _ = ops.attn_fwd(qkv.float(), None, B, S, heads, False)
    got, _, _ = ops.tc_attn_fwd(qkv, None, B, S, heads, False)
    assert norm_err(got, ref) < 4e-3
    ref, _ = ops.attn_fwd(qkv.float(), valid, B, S, heads, True)
    got, _, _ = ops.tc_attn_fwd(qkv, valid, B, S, heads, True)
    assert norm_err(got, ref) < 4e-3
    H = heads * d
    v = qkv.float().view(B, S, 3, heads, d)[:, :, 2]                          # (B,S,heads,d)
    b = int((~valid[:, 0]).nonzero()[0])                                       # a sequence whose first position is padding
    uniform = v[b].mean(dim=0).reshape(H)                                      # query 0 sees no valid key -> mean of all V rows
    torch.testing.assert_close(got.float().view(B, S, H)[b, 0], uniform, rtol=2e-2, atol=2e-2)


@pytest.fixture(params=[(1, 4), (1, 2), (0, 2)], ids=["single_sweep_4wg", "single_sweep_2wg", "two_sweeps"])
def bwd_variant(request, ops):
    """the tensor-core backward kernels: single sweep with four (default) or two epilogue warpgroups, and the older two-sweep one"""
    ops._lib.call("asme_b200_tc_attn_tune", 0, request.param[0])
    ops._lib.call("asme_b200_tc_attn_tune", 1, request.param[1])
    yield request.param
    ops._lib.call("asme_b200_tc_attn_tune", 0, 1)
    ops._lib.call("asme_b200_tc_attn_tune", 1, 4)


@pytest.mark.parametrize("B,S,heads,d", CASES)
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("p_drop", [0.0, 0.2])
def test_tc_attn_bwd_matches_simt(ops, bwd_variant, B, S, heads, d, causal, p_drop):
    """d_qkv of the tensor-core backward vs the fp32 SIMT backward fed with the same (bf16-valued) tensors and the same
    dropout mask; bf16 rounding of P / dS / outputs -> 1e-2 in norm"""
    gen = torch.Generator(device="cuda").manual_seed(B * 77 + S + heads + d)
    qkv, valid = make(gen, B, S, heads, d)
    H = heads * d
    d_ctx = (torch.randn(B * S, H, generator=gen, device="cuda") * 0.5).bfloat16()
    ctx, st, keep = ops.tc_attn_fwd(qkv, valid, B, S, heads, causal, p_drop, 5, 23, save_stats=True)
    got = ops.tc_attn_bwd(qkv, valid, B, S, heads, causal, ctx, d_ctx, st, keep, p_drop)
    if p_drop == 0.0:
        ctx_ref, st_ref = ops.attn_fwd(qkv.float(), valid, B, S, heads, causal, save_stats=True)
        want = ops.attn_bwd(qkv.float(), valid, B, S, heads, causal, ctx_ref, d_ctx.float(), st_ref)
    else:       # autograd of the reference arithmetic under the forward's keep mask
        x = qkv.float().requires_grad_(True)
        torch_attention(x, valid, B, S, heads, causal, unpack_keep(keep, B, heads, S), p_drop).backward(d_ctx.float())
        want = x.grad
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        if S == 1 and name != "dv":
            # a single key: dQ = dK = 0 analytically (dS = P (dP - D) cancels); what is left on either side is the rounding
            # residual of that cancellation -- bounded relative to the gradient scale instead of compared
            assert float(got[:, sl].float().norm()) < 1e-2 * float(want.norm())
            continue
        err = float((got[:, sl].float() - want[:, sl]).norm() / (want[:, sl].norm() + 1e-3 * want.norm()))
        assert err < 1.5e-2, f"{name}: relative norm error {err:.4f}"
    # padded positions beyond each sequence's valid keys receive no dK / dV only through masked keys; rows must be finite
    assert torch.isfinite(got.float()).all()


def test_tc_attn_bwd_fully_masked_rows(ops, bwd_variant):
    gen = torch.Generator(device="cuda").manual_seed(9)
    B, S, heads, d = 3, 70, 2, 32
    qkv, valid = make(gen, B, S, heads, d, pad="left")
    H = heads * d
    d_ctx = (torch.randn(B * S, H, generator=gen, device="cuda") * 0.5).bfloat16()
    ctx_ref, st_ref = ops.attn_fwd(qkv.float(), valid, B, S, heads, True, save_stats=True)
    want = ops.attn_bwd(qkv.float(), valid, B, S, heads, True, ctx_ref, d_ctx.float(), st_ref)
    ctx, st, keep = ops.tc_attn_fwd(qkv, valid, B, S, heads, True, save_stats=True)
    got = ops.tc_attn_bwd(qkv, valid, B, S, heads, True, ctx, d_ctx, st, keep)
    assert norm_err(got, want) < 1.5e-2


@pytest.mark.parametrize("B,S,heads,d", CASES + [(3, 300, 8, 16), (2, 64, 1, 128), (4, 200, 2, 8)])
@pytest.mark.parametrize("causal", [False, True])
def test_attn_row_fwd_matches_reference_arithmetic(ops, B, S, heads, d, causal):
    """one query position per sequence (asme_b200_attn_row_fwd): fp32 restatement of Attention.forward on the same bf16 q/k/v,
    left padding so some query rows see no valid key (uniform attention, quirk Q4); output rounded to bf16 -> 2^-8 relative"""
    gen = torch.Generator(device="cuda").manual_seed(B * 31 + S + heads + d)
    for pad, use_mask in (("right", True), ("left", True), ("right", False)):
        qkv, valid = make(gen, B, S, heads, d, pad=pad)
        kv = valid if use_mask else None
        pos = torch.randint(0, S, (B,), generator=gen, device="cuda")
        pos[0] = 0
        rows = torch.arange(B, device="cuda") * S + pos
        got = ops.attn_row_fwd(qkv, kv, B, S, heads, causal, rows)
        want = torch_attention(qkv, kv, B, S, heads, causal).index_select(0, rows)
        assert got.shape == (B, heads * d) and got.dtype == torch.bfloat16
        torch.testing.assert_close(got.float(), want, rtol=1e-2, atol=1e-2 * float(want.abs().max()))
        assert norm_err(got, want) < 3e-3
        assert torch.equal(got, ops.attn_row_fwd(qkv, kv, B, S, heads, causal, rows))          # deterministic


def test_attn_row_fwd_rejects_unsupported_head_size(ops):
    qkv = torch.zeros(4, 3 * 24, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="head size"):
        ops.attn_row_fwd(qkv, None, 1, 4, 2, False, torch.zeros(1, dtype=torch.long, device="cuda"))


@pytest.mark.parametrize("B,S,heads,d", [(3, 200, 2, 32), (2, 256, 3, 64), (4, 70, 4, 16)])
def test_tc_attn_fwd_rows_is_bit_identical_on_the_selected_rows(ops, B, S, heads, d):
    gen = torch.Generator(device="cuda").manual_seed(S + d)
    qkv, valid = make(gen, B, S, heads, d)
    rows = torch.arange(B, device="cuda") * S + torch.randint(0, S, (B,), generator=gen, device="cuda")
    full, _, _ = ops.tc_attn_fwd(qkv, valid, B, S, heads, True)
    part = ops.tc_attn_fwd_rows(qkv, valid, B, S, heads, True, rows)
    assert torch.equal(part.index_select(0, rows), full.index_select(0, rows))
