"""A training step replayed from its CUDA graph (asme_b200.graphs.GraphedTrainStep: seed / Adam step / lr in device memory)
must be the launch-by-launch step: same kernels, same order, same dropout streams -> identical losses and
weights, bit for bit."""
import pytest
import torch

from test_gpu_models import _random_batch

pytestmark = pytest.mark.gpu


def _setup(seed=0):
    from asme_b200.metrics import build_metrics
    from asme_b200.models import BERT4RecModel
    from asme_b200.modules import MaskedTrainingModule
    torch.manual_seed(seed)
    V, S, H = 503, 40, 64
    model = BERT4RecModel(H, 2, 2, V, S, 0.2)
    module = MaskedTrainingModule(model, metrics=build_metrics({"recall": [10]}), learning_rate=3e-3, num_warmup_steps=1).cuda()
    module.train()
    (opt,), _ = module.configure_optimizers()
    opt.param_groups[0]["lr"] = 3e-3          # the warm-up LambdaLR starts at 0; these tests drive the rate by hand
    return module, opt


def _fresh_batches(n, B, S, V, seed=3):
    """cloze batches with a DIFFERENT number of masked positions each (drawn per batch, as the reference's processor does)"""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        seq, target, _ = _random_batch(gen, B, S, V, p_mask=0.1 + 0.05 * (i % 5))
        out.append({"item": seq.cuda(), "item.target": target.cuda()})
    return out


def test_one_graph_serves_every_batch_and_equals_eager_steps():
    """fresh batch every step, every batch with its own number of target positions: ONE captured graph (the row selection
    and its count live on the device), losses and weights equal to the launch-by-launch steps bit for bit"""
    from asme_b200.graphs import GraphedTrainStep
    V, S, B = 503, 40, 32
    batches = _fresh_batches(12, B, S, V)
    counts = {int((b["item.target"] != 0).sum()) for b in batches}
    assert len(counts) >= 8, "the batches are supposed to differ in their number of masked positions"

    eager, opt_e = _setup()
    losses_e = []
    for i, batch in enumerate(batches):
        opt_e.zero_grad()
        out = eager.training_step(batch, i)
        out["loss"].backward()
        opt_e.step()
        losses_e.append(float(out["loss"].detach()))

    graphed, opt_g = _setup()
    step = GraphedTrainStep(graphed, opt_g, None, warmup_iters=3)      # warm-up steps are undone: the first replay is step 1
    losses_g = [float(step(batch)) for batch in batches]
    assert len(step.graphs) == 1
    assert losses_g == losses_e, (losses_g, losses_e)
    for (n1, p1), (n2, p2) in zip(eager.model.named_parameters(), graphed.model.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2), n1
    assert opt_g._steps == opt_e._steps == len(batches)


def test_graph_replay_takes_the_host_learning_rate():
    from asme_b200.graphs import GraphedTrainStep
    V, S, B = 503, 40, 32
    module, opt = _setup(1)
    step = GraphedTrainStep(module, opt, None)
    batches = _fresh_batches(4, B, S, V, seed=5)
    seen = [float(step(b)) for b in batches]
    assert len(step.graphs) == 1 and len(set(seen)) == 4
    before = [p.detach().clone() for p in module.model.parameters()]
    opt.param_groups[0]["lr"] = 0.0                       # the host-side schedule is written into the device state before each replay
    step(batches[0])
    assert all(torch.equal(a, b) for a, b in zip(before, module.model.parameters()))


def test_no_target_at_all_gives_nan_loss_and_no_update():
    """a batch without any target: nn.CrossEntropyLoss's mean over zero rows is NaN; nothing live reaches the backward"""
    module, opt = _setup(2)
    V, S, B = 503, 40, 8
    seq = torch.randint(3, V, (B, S)).cuda()
    out = module.training_step({"item": seq, "item.target": torch.zeros_like(seq)}, 0)
    assert torch.isnan(out["loss"]).item()
    out["loss"].backward()
    grads = [p.grad for p in module.model.parameters() if p.grad is not None]
    assert grads and all(torch.count_nonzero(g) == 0 for g in grads)


def test_validation_keeps_the_optimizer_state():
    """Trainer.validate must not re-pack the parameter arena (it used to: module.to() on every pass dropped Adam's moments)"""
    from asme_b200.trainer import Trainer
    module, _ = _setup(3)
    V, S, B = 503, 40, 16
    train = _fresh_batches(3, B, S, V, seed=7)
    g = torch.Generator().manual_seed(11)
    val = []
    for _ in range(2):
        seq = torch.randint(3, V, (B, S), generator=g)
        seq[:, -1] = 1
        val.append({"item": seq.cuda(), "item.target": torch.randint(3, V, (B,), generator=g).cuda()})
    trainer = Trainer(max_epochs=2, device=torch.device("cuda"), gradient_clip_val=5.0)
    arena = module.model._arena
    ptr = arena.flat.data_ptr()
    trainer.fit(module, train, val)
    assert arena.flat.data_ptr() == ptr and arena.exp_avg is not None and float(arena.exp_avg.abs().sum()) > 0
    assert "recall@10" in trainer.history[-1]


def test_graphed_eval_step_equals_eager():
    """GraphedEvalStep replays model.evaluate_rank: same top-k, ranks and target scores as the launch-by-launch step, for new
    inputs copied into the captured buffers"""
    from asme_b200.graphs import GraphedEvalStep
    from asme_b200.models import BERT4RecModel
    torch.manual_seed(0)
    V, S, H, B = 5003, 40, 64, 96
    model = BERT4RecModel(H, 2, 2, V, S, 0.1, initializer_range=0.2).cuda().eval()

    def batch(seed):
        g = torch.Generator().manual_seed(seed)
        seq = torch.randint(3, V, (B, S), generator=g)
        lengths = torch.randint(2, S, (B,), generator=g)
        seq = torch.where(torch.arange(S).unsqueeze(0) < lengths.unsqueeze(1), seq, torch.zeros_like(seq))
        seq[torch.arange(B), lengths] = 1
        return {"seq": seq.cuda(), "target": torch.randint(3, V, (B,), generator=g).cuda()}

    def fn(b):
        return model.evaluate_rank(b["seq"], b["seq"].ne(0), {}, b["target"], k=10, full_rank=True)

    graphed = GraphedEvalStep(fn)
    for seed in (1, 2, 3):
        b = batch(seed)
        got = {k: v.clone() for k, v in graphed(b).items()}
        want = fn(b)
        assert len(graphed.graphs) == 1
        for k in ("topk_idx", "topk_val", "rank", "target_score"):
            assert torch.equal(got[k], want[k]), k


def test_module_eval_graph_option_gives_the_same_metrics():
    """module.eval_graph = True: validation_step replays the captured model part; step values and epoch metrics are equal to the
    launch-by-launch hooks"""
    from asme_b200.metrics import build_metrics
    from asme_b200.models import BERT4RecModel, SASRecModel
    from asme_b200.modules import MaskedTrainingModule, NextItemPredictionTrainingModule
    torch.manual_seed(0)
    V, S, H, B = 2003, 30, 64, 64

    class Tok:
        pad_token_id, mask_token_id = 0, 1

    def batches(masked):
        out = []
        for seed in (1, 2, 3):
            g = torch.Generator().manual_seed(seed)
            seq = torch.randint(3, V, (B, S), generator=g)
            lengths = torch.randint(2, S, (B,), generator=g)
            seq = torch.where(torch.arange(S).unsqueeze(0) < lengths.unsqueeze(1), seq, torch.zeros_like(seq))
            if masked:
                seq[torch.arange(B), lengths] = 1
            out.append({"item": seq.cuda(), "item.target": torch.randint(3, V, (B,), generator=g).cuda()})
        return out

    for cls, mod_cls, kw, masked in ((BERT4RecModel, MaskedTrainingModule, {"initializer_range": 0.2}, True),
                                     (SASRecModel, NextItemPredictionTrainingModule, {"mode": "full"}, False)):
        model = cls(H, 2, 2, V, S, 0.1, **kw)
        results = []
        for use_graph in (False, True):
            module = mod_cls(model, item_tokenizer=Tok(), metrics=build_metrics({"recall": [1, 10], "ndcg": [10], "mrr": [10], "rank": []})).cuda().eval()
            module.eval_graph = use_graph
            steps = [module.validation_step_end(module.validation_step(b, i)) for i, b in enumerate(batches(masked))]
            results.append((steps, module.validation_epoch_end(None)))
            if use_graph:
                assert len(module._eval_graphs) == 1
        (s0, e0), (s1, e1) = results
        for a, b in zip(s0, s1):
            assert set(a) == set(b) and all(torch.equal(a[k].float().cpu(), b[k].float().cpu()) for k in a)
        assert all(torch.equal(e0[k].cpu(), e1[k].cpu()) for k in e0)


def test_second_stream_leaves_do_not_change_a_single_bit(monkeypatch):
    """the leaves of the backward pass that run on the second stream / as parallel graph branches -- encoder weight gradients, the
    catalog-gradient sweep, the id sort, the (gamma, beta) reductions of every LayerNorm backward, the modifier's weight gradient --
    are the same kernels with the same order of additions per destination: losses and weights after several steps equal the
    one-stream run bit for bit"""
    from asme_b200 import engine
    V, S, B = 503, 40, 32
    batches = _fresh_batches(6, B, S, V, seed=11)
    runs = {}
    for side in (True, False):
        monkeypatch.setattr(engine, "SIDE_STREAM_WGRAD", side)
        module, opt = _setup(seed=5)
        losses = []
        for i, batch in enumerate(batches):
            opt.zero_grad()
            out = module.training_step(batch, i)
            out["loss"].backward()
            opt.step()
            losses.append(float(out["loss"].detach()))
        torch.cuda.synchronize()
        runs[side] = (losses, {n: p.detach().clone() for n, p in module.model.named_parameters()})
    assert runs[True][0] == runs[False][0]
    for n, p in runs[True][1].items():
        assert torch.equal(p, runs[False][1][n]), n
