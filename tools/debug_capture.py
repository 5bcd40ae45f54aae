import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")]
import torch
from asme_b200 import ops
from asme_b200.models import BERT4RecModel
from asme_b200.modules import MaskedTrainingModule
from asme_b200.metrics import build_metrics
from asme_b200.graphs import StepState
sys.path.insert(0, ROOT)
import bench

dev = torch.device("cuda")
torch.manual_seed(0)
cfg = bench.C2
model = BERT4RecModel(cfg["H"], cfg["heads"], cfg["L"], cfg["V"], cfg["S"], cfg["dropout"])
module = MaskedTrainingModule(model, metrics=build_metrics({"recall": [10]}), learning_rate=1e-3, num_warmup_steps=100).to(dev)
module.train()
(optimizer,), _ = module.configure_optimizers()
gen = torch.Generator().manual_seed(1)
inp, target, rows = bench.make_cloze_batch(gen, 64, cfg["S"], cfg["V"], 0.2, 20)
batch = {"item": inp.to(dev), "item.target": target.to(dev)}


def attempt(name, fn, mode="global"):
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        g.replay()
        torch.cuda.synchronize()
        print(f"OK    {name} [{mode}]")
    except Exception as e:
        print(f"FAIL  {name} [{mode}]: {str(e).splitlines()[0][:150]}")
        torch.cuda.synchronize()


state = StepState(dev, seed=5, adam_step=1, lr=1e-3)
x = torch.randn(51200 // 4, 64, device=dev)
x16 = x.bfloat16()
w16 = torch.randn(192, 64, device=dev).bfloat16()
attempt("step_state_advance", lambda: ops.step_state_advance(state.tensor))
attempt("cast_bf16", lambda: ops.cast_bf16(x))
attempt("tc_gemm", lambda: ops.tc_gemm(x16, w16, out_f32=False, out_bf16=True))
qkv = ops.tc_gemm(x16, w16, out_f32=False, out_bf16=True)["bf16"]
attempt("tc_attn_fwd", lambda: ops.tc_attn_fwd(qkv, None, 64, 200, 2, False, 0.2, state.indirect_seed(), 3, save_stats=True))
ids = torch.randint(0, 3709, (12800,), device=dev)
tab = torch.zeros(3709, 64, device=dev)
attempt("embgrad", lambda: ops.embgrad_sorted_reduce(ids, x, tab))
attempt("layernorm_bwd", lambda: ops.layernorm_bwd(x, x, torch.ones(64, device=dev), torch.ones(2, x.shape[0], device=dev), torch.zeros(2, 64, device=dev)))
attempt("forward loss_ce", lambda: model.loss_ce(batch["item"], batch["item"].ne(0), {}, batch["item.target"]))


def fwd_bwd():
    loss, ctx = model.loss_ce(batch["item"], batch["item"].ne(0), {}, batch["item.target"])
    model.loss_ce_backward(ctx)


attempt("fwd + fused backward (no autograd)", fwd_bwd)


def step_autograd():
    optimizer.zero_grad()
    out = module.training_step(batch, 0)
    out["loss"].backward()


attempt("training_step + loss.backward()", step_autograd)
attempt("training_step + loss.backward()", step_autograd, "thread_local")


def full():
    optimizer.zero_grad()
    out = module.training_step(batch, 0)
    out["loss"].backward()
    optimizer.step()


attempt("full step", full)
attempt("full step", full, "thread_local")
