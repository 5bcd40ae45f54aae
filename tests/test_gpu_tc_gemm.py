"""Parity tests of the tensor-core dense layers (asme_b200_tc_gemm / asme_b200_tc_wgrad) against float64 restatements
of the reference's nn.Linear arithmetic (transformer_layers.py:190-220) on the same bf16-rounded operands.
Small-integer operands make every product / partial sum exact, so results must match bit for bit whatever the
accumulation order -- this pins the K-major and MN-major shared-memory descriptors."""
import math

import pytest
import torch

from oracle import asme_oracle as O

pytestmark = pytest.mark.gpu

SHAPES = [(1, 64, 64), (100, 192, 64), (128, 256, 64), (300, 64, 256), (257, 128, 128), (130, 96, 192), (51200, 192, 64),
          (300, 384, 128), (200, 512, 128), (200, 128, 512), (77, 640, 320)]


@pytest.fixture(scope="module")
def ops():
    from asme_b200 import ops
    return ops


def ints(gen, *shape, lo=-3, hi=4):
    return torch.randint(lo, hi, shape, generator=gen, device="cuda").float()


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_forward_integer_exact(ops, M, N, K):
    gen = torch.Generator(device="cuda").manual_seed(M + N + K)
    a, w, bias = ints(gen, M, K), ints(gen, N, K), ints(gen, N)
    out = ops.tc_gemm(a.bfloat16(), w.bfloat16(), bias=bias)
    assert torch.equal(out["f32"].double(), a.double() @ w.double().t() + bias.double())


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_dgrad_mn_major_integer_exact(ops, M, N, K):
    """C = A B with B (K,N) row-major: the MN-major B descriptor (dX = dY W without a transposed copy of W)"""
    gen = torch.Generator(device="cuda").manual_seed(M + N + 2 * K)
    a, b = ints(gen, M, K), ints(gen, K, N)
    out = ops.tc_gemm(a.bfloat16(), b.bfloat16(), b_is_kn=True)
    assert torch.equal(out["f32"].double(), a.double() @ b.double())


@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (100, 192, 64), (1000, 256, 64), (5000, 64, 256), (333, 128, 192), (51200, 256, 64),
                                   (700, 384, 128), (700, 512, 128), (700, 128, 512)])
def test_tc_wgrad_integer_exact(ops, M, N, K):
    gen = torch.Generator(device="cuda").manual_seed(M + 3 * N + K)
    dy, x = ints(gen, M, N, lo=-2, hi=3), ints(gen, M, K, lo=-2, hi=3)
    dw = ints(gen, N, K)
    db = ints(gen, N)
    want_w = dw.double() + dy.double().t() @ x.double()
    want_b = db.double() + dy.double().sum(0)
    ops.tc_wgrad(dy.bfloat16(), x.bfloat16(), dw, db, accumulate=True)
    assert torch.equal(dw.double(), want_w)
    assert torch.equal(db.double(), want_b)
    ops.tc_wgrad(dy.bfloat16(), x.bfloat16(), dw, db, accumulate=False)
    assert torch.equal(dw.double(), dy.double().t() @ x.double())


def test_tc_gemm_epilogue_matches_oracle(ops):
    """bias -> (pre-activation copy) -> exact-erf GELU -> residual, fp32 + bf16 outputs; gaussian operands"""
    gen = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 1000, 256, 64
    a = torch.randn(M, K, generator=gen, device="cuda").bfloat16()
    w = (torch.randn(N, K, generator=gen, device="cuda") * 0.2).bfloat16()
    bias = torch.randn(N, generator=gen, device="cuda")
    res = torch.randn(M, N, generator=gen, device="cuda")
    out = ops.tc_gemm(a, w, bias=bias, act=1, residual=res, out_bf16=True, pre_act=True)
    z = a.double() @ w.double().t() + bias.double()
    want = O.gelu_erf(z) + res.double()
    torch.testing.assert_close(out["f32"].double(), want, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out["pre"].double(), z, rtol=2 ** -8, atol=1e-6)        # bf16 rounding of the stored copy
    torch.testing.assert_close(out["bf16"].double(), want, rtol=2 ** -8, atol=1e-6)
    # backward through GELU: (dy W) * gelu'(z)
    dy = torch.randn(M, N, generator=gen, device="cuda").bfloat16()
    w2 = (torch.randn(N, K, generator=gen, device="cuda") * 0.2).bfloat16()              # (N_out=N, K_in=K): dX = dY W
    zz = torch.randn(M, K, generator=gen, device="cuda").bfloat16()
    got = ops.tc_gemm(dy, w2, b_is_kn=True, gelu_grad_of=zz)["f32"]
    zd = zz.double()
    grad = 0.5 * (1 + torch.erf(zd / math.sqrt(2))) + zd * torch.exp(-0.5 * zd * zd) / math.sqrt(2 * math.pi)
    torch.testing.assert_close(got.double(), (dy.double() @ w2.double()) * grad, rtol=1e-5, atol=1e-5)
    # bf16-only outputs (what the encoder's feed-forward uses) take the 7-instruction GELU: x sigmoid(2u) with a fitted quintic u,
    # |error| <= 2.6e-5 absolute against the erf form (1.1e-4 for the derivative) before the bf16 rounding of the store
    out16 = ops.tc_gemm(a, w, bias=bias, act=1, out_f32=False, out_bf16=True)["bf16"]
    torch.testing.assert_close(out16.double(), O.gelu_erf(z), rtol=2 ** -8, atol=4e-5)
    got16 = ops.tc_gemm(dy, w2, b_is_kn=True, gelu_grad_of=zz, out_f32=False, out_bf16=True)["bf16"]
    base = dy.double() @ w2.double()
    torch.testing.assert_close(got16.double(), base * grad, rtol=2 ** -8, atol=2e-4 * float(base.abs().max()))


def test_tc_gemm_dropout_uses_the_shared_stream(ops):
    """the fused epilogue dropout must produce exactly the mask of the stand-alone dropout kernel (same seed / site /
    element index), because the backward pass re-creates masks instead of storing them"""
    gen = torch.Generator(device="cuda").manual_seed(2)
    M, N, K = 700, 192, 64
    a, w = ints(gen, M, K), ints(gen, N, K)
    plain = ops.tc_gemm(a.bfloat16(), w.bfloat16())["f32"]
    dropped = ops.tc_gemm(a.bfloat16(), w.bfloat16(), p_drop=0.25, seed=1234567, site=21)["f32"]
    want = ops.dropout(plain, 0.25, 1234567, 21)
    assert torch.equal(dropped, want)
    keep = (want != 0).float().mean().item()
    assert abs(keep - 0.75 * (plain != 0).float().mean().item()) < 0.02


@pytest.fixture(params=[1, 0], ids=["persistent", "per_tile"])
def gemm_variant(request, ops):
    """both tall kernels: persistent CTAs with a resident weight tile (default) and one CTA per tile"""
    ops._lib.call("asme_b200_tc_gemm_tune", 0, request.param)
    yield request.param
    ops._lib.call("asme_b200_tc_gemm_tune", 0, 1)


@pytest.mark.parametrize("M,N,K", [(100, 192, 64), (51200, 64, 256), (300, 384, 128), (200, 128, 512), (1, 64, 64)])
def test_tc_gemm_variants_agree(ops, gemm_variant, M, N, K):
    gen = torch.Generator(device="cuda").manual_seed(M + N + K)
    a, w, bias = ints(gen, M, K), ints(gen, N, K), ints(gen, N)
    out = ops.tc_gemm(a.bfloat16(), w.bfloat16(), bias=bias)
    assert torch.equal(out["f32"].double(), a.double() @ w.double().t() + bias.double())


@pytest.mark.parametrize("M,N,K", [(300, 64, 64), (51200, 64, 256), (257, 128, 512), (130, 32, 64), (1000, 128, 192)])
def test_tc_gemm_fused_layernorm(ops, M, N, K):
    """LayerNorm fused into the epilogue (transformer_layers.py:120-130, the next sublayer's norm) against the oracle's
    LayerNorm of the very same fp32 rows, and against the stand-alone LayerNorm kernel"""
    gen = torch.Generator(device="cuda").manual_seed(M * 3 + N + K)
    a = (torch.randn(M, K, generator=gen, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, generator=gen, device="cuda") * 0.2).bfloat16()
    bias = torch.randn(N, generator=gen, device="cuda")
    res = torch.randn(M, N, generator=gen, device="cuda") * 2 + 0.5
    gamma = torch.randn(N, generator=gen, device="cuda") * 0.3 + 1.0
    beta = torch.randn(N, generator=gen, device="cuda") * 0.2
    old = ops.FUSE_LAYERNORM
    try:
        ops.FUSE_LAYERNORM = True
        fused = ops.tc_gemm(a, w, bias=bias, residual=res, ln=(gamma, beta), ln_stats=True)
        ops.FUSE_LAYERNORM = False
        split = ops.tc_gemm(a, w, bias=bias, residual=res, ln=(gamma, beta), ln_stats=True)
    finally:
        ops.FUSE_LAYERNORM = old
    assert torch.equal(fused["f32"], split["f32"])
    want = O.layer_norm(fused["f32"].double().cpu(), gamma.double().cpu(), beta.double().cpu())
    torch.testing.assert_close(fused["ln16"].double().cpu(), want, rtol=1e-2, atol=1e-2)             # bf16 output
    x = fused["f32"].double()
    torch.testing.assert_close(fused["ln_st"][0].double(), x.mean(dim=1), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(fused["ln_st"][1].double(), 1.0 / torch.sqrt(x.var(dim=1, unbiased=False) + 1e-5), rtol=2e-5, atol=0)
    # the two implementations round to the same bf16 value almost everywhere; where they differ it is by one bf16 ulp
    diff = (fused["ln16"].float() - split["ln16"].float()).abs()
    assert float(diff.max()) <= 2.0 ** -7 * float(split["ln16"].float().abs().max())
    assert float((diff > 0).float().mean()) < 0.02
    torch.testing.assert_close(fused["ln_st"], split["ln_st"], rtol=2e-5, atol=1e-6)


# ---- feed-forward block as one kernel (asme_b200_tc_ffn_fused, evaluation path) ----------------------------------------------
FFN_SHAPES = [(1, 64, 256), (300, 128, 512), (1000, 64, 256), (1024, 128, 512), (148 * 128 * 2 + 77, 128, 512), (148 * 128 + 130, 64, 128),
              (700, 128, 192)]


@pytest.mark.parametrize("M,H,FF", FFN_SHAPES)
def test_tc_ffn_fused_equals_two_gemms(ops, M, H, FF):
    """out = x + W2 gelu(W1 y + b1) + b2 (transformer_layers.py:217-220, :120-130): the fused kernel against the two tensor-core
    GEMM launches it replaces -- bit for bit on the fp32 rows -- and its fused LayerNorm against the stand-alone LayerNorm kernel on
    the same rows (one bf16 ulp: the row statistics are summed in a different order)."""
    from asme_b200._lib import ACT_GELU
    gen = torch.Generator(device="cuda").manual_seed(M + H + FF)
    y = (torch.randn(M, H, generator=gen, device="cuda")).bfloat16()
    x = torch.randn(M, H, generator=gen, device="cuda")
    w1 = (torch.randn(FF, H, generator=gen, device="cuda") * 0.1).bfloat16()
    w2 = (torch.randn(H, FF, generator=gen, device="cuda") * 0.1).bfloat16()
    b1 = torch.randn(FF, generator=gen, device="cuda") * 0.1
    b2 = torch.randn(H, generator=gen, device="cuda") * 0.1
    g = 1.0 + 0.1 * torch.randn(H, generator=gen, device="cuda")
    b = 0.1 * torch.randn(H, generator=gen, device="cuda")
    a16 = ops.tc_gemm(y, w1, bias=b1, act=ACT_GELU, out_f32=False, out_bf16=True)["bf16"]
    want = ops.tc_gemm(a16, w2, bias=b2, residual=x)["f32"]
    want_ln, _, _ = ops.layernorm_fwd_bf16(want, g, b)
    got = ops.tc_ffn_fused(y, w1, b1, w2, b2, x, ln=(g, b))
    torch.cuda.synchronize()
    assert torch.equal(got["f32"], want)
    d = (got["ln16"].float() - want_ln.float()).abs()
    ulp = want_ln.float().abs().clamp_min(2.0 ** -6) * 2.0 ** -7
    assert bool((d <= ulp).all()), float((d / ulp).max())
    assert float((d > 0).float().mean()) < 0.01
    # the oracle's arithmetic on the same bf16 operands (float64): gelu = 0.5 z (1 + erf(z / sqrt 2))
    z = y.double() @ w1.double().t() + b1.double()
    a = (0.5 * z * (1.0 + torch.erf(z / math.sqrt(2.0)))).float().bfloat16().double()
    ref = x.double() + a @ w2.double().t() + b2.double()
    assert float((got["f32"].double() - ref).abs().max()) < 2e-2 * float(ref.abs().max())
    only = ops.tc_ffn_fused(y, w1, b1, w2, b2, x, ln=None)
    assert only["ln16"] is None and torch.equal(only["f32"], want)
    ln_only = ops.tc_ffn_fused(y, w1, b1, w2, b2, x, ln=(g, b), out_f32=False)
    assert ln_only["f32"] is None and torch.equal(ln_only["ln16"], got["ln16"])


def test_tc_ffn_fused_integer_exact(ops):
    """small-integer operands and a zero first-layer bias keep z in {integers}: with gelu(0) = 0 and W1 = 0 the block reduces to
    out = x + b2 exactly; with identity-like W1 rows it pins the chunk order of the second contraction"""
    gen = torch.Generator(device="cuda").manual_seed(5)
    M, H, FF = 260, 128, 512
    y = ints(gen, M, H).bfloat16()
    x = ints(gen, M, H)
    w2 = ints(gen, H, FF).bfloat16()
    b2 = ints(gen, H)
    out = ops.tc_ffn_fused(y, torch.zeros(FF, H, device="cuda").bfloat16(), torch.zeros(FF, device="cuda"), w2, b2, x)["f32"]
    assert torch.equal(out, x + b2)


@pytest.mark.parametrize("M,H,FF", [(1, 64, 256), (300, 128, 512), (1024, 128, 512), (148 * 128 * 2 + 77, 128, 512), (148 * 128 * 3 + 5, 64, 256),
                                    (700, 128, 64)])
def test_tc_block_tail_fused_equals_the_unfused_tail(ops, M, H, FF):
    """output projection + residual + LayerNorm + feed-forward + residual (+ next LayerNorm) in one kernel (transformer_layers.py:181-199,
    :120-130, :217-220, :251-258) against the launches it replaces.  x2 = x + ctx Wo^T + bo is the same arithmetic (checked through a
    zero feed-forward); the LayerNorm inside sums its statistics in another order than the stand-alone kernel, so y and everything
    after it agree to bf16 rounding, not bit for bit."""
    from asme_b200._lib import ACT_GELU
    gen = torch.Generator(device="cuda").manual_seed(M + H + FF + 7)
    ctx = torch.randn(M, H, generator=gen, device="cuda").bfloat16()
    x = torch.randn(M, H, generator=gen, device="cuda")
    wo = (torch.randn(H, H, generator=gen, device="cuda") * 0.1).bfloat16()
    bo = torch.randn(H, generator=gen, device="cuda") * 0.1
    w1 = (torch.randn(FF, H, generator=gen, device="cuda") * 0.1).bfloat16()
    w2 = (torch.randn(H, FF, generator=gen, device="cuda") * 0.1).bfloat16()
    b1 = torch.randn(FF, generator=gen, device="cuda") * 0.1
    b2 = torch.randn(H, generator=gen, device="cuda") * 0.1
    g2, be2 = 1.0 + 0.1 * torch.randn(H, generator=gen, device="cuda"), 0.1 * torch.randn(H, generator=gen, device="cuda")
    g3, be3 = 1.0 + 0.1 * torch.randn(H, generator=gen, device="cuda"), 0.1 * torch.randn(H, generator=gen, device="cuda")
    x2 = ops.tc_gemm(ctx, wo, bias=bo, residual=x)["f32"]
    y2, _, _ = ops.layernorm_fwd_bf16(x2, g2, be2)
    a16 = ops.tc_gemm(y2, w1, bias=b1, act=ACT_GELU, out_f32=False, out_bf16=True)["bf16"]
    want = ops.tc_gemm(a16, w2, bias=b2, residual=x2)["f32"]
    want_ln, _, _ = ops.layernorm_fwd_bf16(want, g3, be3)
    # (1) zero feed-forward weights: out = x2 + b2 exactly -> the prologue's arithmetic and the TMEM round trip of x2 are bit-exact
    z1, z2 = torch.zeros_like(w1), torch.zeros_like(w2)
    only_x2 = ops.tc_block_tail_fused(ctx, wo, bo, x, (g2, be2), z1, torch.zeros_like(b1), z2, b2)["f32"]
    assert torch.equal(only_x2, x2 + b2)
    # (2) the whole tail
    got = ops.tc_block_tail_fused(ctx, wo, bo, x, (g2, be2), w1, b1, w2, b2, ln=(g3, be3))
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    assert float((got["f32"] - want).abs().max()) < 2e-2 * scale
    assert float((got["f32"] - want).norm() / want.norm()) < 2e-3
    assert float((got["ln16"].float() - want_ln.float()).norm() / want_ln.float().norm()) < 5e-3
    # float64 restatement on the same bf16 operands, LayerNorm output rounded to bf16 where the kernels round it
    x2d = x.double() + ctx.double() @ wo.double().t() + bo.double()
    yd = torch.nn.functional.layer_norm(x2d, (H,), g2.double(), be2.double(), 1e-5).float().bfloat16().double()
    zd = yd @ w1.double().t() + b1.double()
    ad = (0.5 * zd * (1.0 + torch.erf(zd / math.sqrt(2.0)))).float().bfloat16().double()
    ref = x2d + ad @ w2.double().t() + b2.double()
    assert float((got["f32"].double() - ref).norm() / ref.norm()) < 2e-3
    again = ops.tc_block_tail_fused(ctx, wo, bo, x, (g2, be2), w1, b1, w2, b2, ln=(g3, be3))
    assert torch.equal(again["f32"], got["f32"]) and torch.equal(again["ln16"], got["ln16"])      # deterministic
