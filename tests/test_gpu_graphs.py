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


def test_graph_replay_equals_eager_steps():
    from asme_b200.graphs import GraphedTrainStep
    V, S, B = 503, 40, 32
    seq, target, _ = _random_batch(torch.Generator().manual_seed(3), B, S, V)
    rows = torch.nonzero(target.reshape(-1) != 0).reshape(-1)
    batch = {"item": seq.cuda(), "item.target": target.cuda(), "_target_rows": rows.cuda()}
    n_warm, n_steps = 3, 6

    eager, opt_e = _setup()
    losses_e = []
    for i in range(n_warm + n_steps):
        opt_e.zero_grad()
        out = eager.training_step(batch, i)
        out["loss"].backward()
        opt_e.step()
        losses_e.append(float(out["loss"].detach()))

    graphed, opt_g = _setup()
    step = GraphedTrainStep(graphed, opt_g, None, warmup_iters=n_warm)      # capture runs n_warm real steps on the batch first
    losses_g = [float(step(batch)) for _ in range(n_steps)]
    assert losses_g == losses_e[n_warm:], (losses_g, losses_e)
    for (n1, p1), (n2, p2) in zip(eager.model.named_parameters(), graphed.model.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2), n1
    assert losses_g[-1] < losses_g[0]


def test_graph_replay_takes_new_inputs_and_learning_rate():
    from asme_b200.graphs import GraphedTrainStep
    V, S, B = 503, 40, 32
    module, opt = _setup(1)
    step = GraphedTrainStep(module, opt, None)
    gen = torch.Generator().manual_seed(5)
    seen = []
    for i in range(4):
        seq, target, _ = _random_batch(gen, B, S, V)
        rows = torch.nonzero(target.reshape(-1) != 0).reshape(-1)
        rows = torch.nn.functional.pad(rows, (0, 400 - rows.numel()), value=int(rows[-1]))[:400] if rows.numel() < 400 else rows[:400]
        # same signature (shapes) -> the same graph is replayed with new data copied into its static inputs
        batch = {"item": seq.cuda(), "item.target": target.cuda(), "_target_rows": rows.cuda()}
        seen.append(float(step(batch)))
    assert len(step.graphs) == 1 and len(set(seen)) == 4
    before = [p.detach().clone() for p in module.model.parameters()]
    opt.param_groups[0]["lr"] = 0.0                       # the host-side schedule is written into the device state before each replay
    step(batch)
    assert all(torch.equal(a, b) for a, b in zip(before, module.model.parameters()))


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
