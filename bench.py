#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native ASME hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

BASELINE.json's metric has two halves, both measured here:

  train (C2, the headline ``value``): BERT4Rec cloze TRAINING step at the ML-1M shape -- 3,706 items (+3 special tokens, V=3709),
      max_seq_len 200, hidden 64, 2 layers, 2 heads, batch 256 per GPU, dropout 0.2, Adam (0.99/0.998) with warm-up schedule:
      forward + fused scoring/cross-entropy + backward + optimizer step.  64 DISTINCT synthetic batches (uniform ids, lengths
      U[20,200], right padded), cloze-masked on the GPU by the input pipeline (every batch has its own number of masked positions);
      ONE CUDA graph serves all of them.
  eval  (C5, the ``eval`` object -- last key of the line): full-catalog scoring + exact top-10 + Recall/NDCG@10 on a synthetic
      1M-item catalog, hidden 128, seq 200, 1024 users per GPU per step; with N GPUs the catalog is vocab-sharded.

One JSON line on stdout (rank 0):
  value / ms_per_step   train seqs/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e                   the same through the public module API with HOST (pinned) inputs: H2D copy of the step's item sequences,
                        GPU cloze masking, the step, D2H read of the loss -- all inside the timed region, every step; in the
                        training loop the copies are double-buffered as a prefetching loader would (H2D of step i+1 on a copy
                        stream while step i runs, the loss of step i read one step later; the evaluation loop, whose step_end
                        hands metric values to the host, stays synchronous: pipelined the same way it was slower, 1.195 vs 1.145 ms
                        per step in one run -- measured while the metrics update still synchronised the stream every step
                        (asme_b200/metrics.py, fixed since); not re-measured after that fix)
  roofline              dominant kernel of the step: algorithmic bytes (or flops) per launch / CUDA-event launch duration
  cpu_baseline          the CPU oracle (a port of the reference's PyTorch path) on this box's host cores, full C2 batch
  eval                  {value users/s, e2e, roofline (scoring sweep, flops on H=128), cpu_baseline (B=64), recall@10, checks}
The per-kernel tables go to stderr.  --impl reference times the CPU oracle port with all host threads on the same workloads.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")
for _p in (ROOT, PKG_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

C2 = dict(V=3709, S=200, H=64, L=2, heads=2, B=256, dropout=0.2, mask_prob=0.2, only_last_prob=0.1, min_len=20)
C5 = dict(V=1_000_003, S=200, H=128, L=2, heads=2, B=1024, k=10, B_cpu=64)
METRIC = "train_seqs_per_sec"
UNIT = "seqs/s"
N_BATCHES = 64


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def workload_config(world):
    """the workload -- identical for this arm and the reference arm; how each arm runs it is in its ``implementation`` key"""
    cfg = C2
    return {"workload": "C2 BERT4Rec cloze training, ML-1M shape", "items": cfg["V"], "seq_len": cfg["S"], "hidden": cfg["H"],
            "layers": cfg["L"], "heads": cfg["heads"], "batch_per_gpu": cfg["B"], "global_batch": cfg["B"] * world, "dropout": cfg["dropout"],
            "optimizer": "Adam(0.99,0.998), LambdaLR warm-up", "parallelism": f"dp{world}" if world > 1 else "single",
            "l2": "inputs larger than L2: each step streams ~0.9 GB of activations (> 126 MB L2); distinct batches rotate"}


# ----------------------------------------------------------------------------------------------------------------
# synthetic data (seeded as SURVEY.md 8d: manual_seed(1234 + config index))
# ----------------------------------------------------------------------------------------------------------------
def make_sessions(gen, B, S, V, min_len):
    """right-padded item sequences as the reference's collate delivers them (before any masking)"""
    seq = torch.randint(3, V, (B, S), generator=gen)
    lengths = torch.randint(min_len, S + 1, (B,), generator=gen)
    return torch.where(torch.arange(S).unsqueeze(0) < lengths.unsqueeze(1), seq, torch.zeros_like(seq))


def cloze_host(gen, seq, mask_prob, only_last_prob):
    """CPU cloze masking of one batch for the oracle arms (same distribution as the GPU input pipeline)"""
    valid = seq.ne(0)
    lengths = valid.sum(1)
    B = seq.shape[0]
    masked = (torch.rand(seq.shape, generator=gen) < mask_prob) & valid
    last = torch.zeros_like(masked)
    last[torch.arange(B), lengths - 1] = True
    only_last = torch.rand(B, generator=gen) < only_last_prob
    masked = torch.where(only_last.unsqueeze(1), last, masked | (last & ~masked.any(dim=1, keepdim=True)))
    return torch.where(masked, torch.ones_like(seq), seq), torch.where(masked, seq, torch.zeros_like(seq))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def count_in(self, windows):
        return sum(1 for ts, _ in self.samples if any(a <= ts <= b for a, b in windows))

    def summary(self, windows):
        sm, mx, reasons = [], [], set()
        for ts, line in self.samples:
            if not any(a <= ts <= b for a, b in windows):
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU oracle steps (cpu_baseline and --impl reference): the oracle is the checker / the reported baseline, never the product
# ----------------------------------------------------------------------------------------------------------------
def oracle_train_setup(cfg, B_cpu, seed=0):
    from oracle import asme_oracle as O
    from asme_b200.models import BERT4RecModel
    torch.manual_seed(seed)
    model = BERT4RecModel(cfg["H"], cfg["heads"], cfg["L"], cfg["V"], cfg["S"], cfg["dropout"])
    w = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items() if k != "_projection_layer.embedding.weight"}
    names = list(w)
    m = [torch.zeros_like(w[n]) for n in names]
    v = [torch.zeros_like(w[n]) for n in names]
    gen = torch.Generator().manual_seed(1234 + 1)
    batches = [cloze_host(gen, make_sessions(gen, B_cpu, cfg["S"], cfg["V"], cfg["min_len"]), cfg["mask_prob"], cfg["only_last_prob"])
               for _ in range(4)]
    state = dict(step=0)

    def step():
        inp, target = batches[state["step"] % len(batches)]
        O.DROPOUT_P = cfg["dropout"]
        for n in names:
            w[n].grad = None
        loss = O.cross_entropy_ignore_pad(O.bert4rec_logits(w, inp, cfg["heads"], cfg["L"]), target)
        loss.backward()
        state["step"] += 1
        with torch.no_grad():
            O.adam_step([w[n] for n in names], [w[n].grad for n in names], m, v, state["step"], 1e-3, 0.99, 0.998, 1e-8, 0.0)
        O.DROPOUT_P = 0.0
        return float(loss.detach())

    return step


def oracle_eval_setup(cfg, B_cpu, seed=0):
    """full-catalog evaluation as the reference runs it, restricted to what a CPU can hold: encoder on B_cpu users, MASK-row select,
    modifier, projection of the SELECTED rows onto the catalog (the reference projects every position: (B,S,V) = 51 GB at this
    shape -- the port is given the cheaper form), then the reference's metric path: one full argsort per metric (common.py:18-27)"""
    from oracle import asme_oracle as O
    from asme_b200.models import BERT4RecModel
    torch.manual_seed(seed)
    model = BERT4RecModel(cfg["H"], cfg["heads"], cfg["L"], cfg["V"], cfg["S"], 0.0)
    w = {k: v.detach().clone() for k, v in model.state_dict().items() if k != "_projection_layer.embedding.weight"}
    gen = torch.Generator().manual_seed(1234 + 4)
    seq = make_sessions(gen, B_cpu, cfg["S"] - 1, cfg["V"], 20)
    seq = torch.cat([seq, torch.zeros(B_cpu, 1, dtype=seq.dtype)], dim=1)
    seq[torch.arange(B_cpu), seq.ne(0).sum(1)] = 1
    target = torch.randint(3, cfg["V"], (B_cpu,), generator=gen)
    table, bias = w["_sequence_embedding_layer.item_embedding.embedding.weight"], w["_projection_layer.output_bias"]

    def step():
        with torch.no_grad():
            hidden = O.bert4rec_hidden(w, seq, cfg["heads"], cfg["L"])
            rows = O.ffn_modifier(O.select_masked_rows(hidden, seq), w)
            logits = O.project(rows, table, bias)
            out = {}
            for name in ("recall", "ndcg"):                       # one sort per metric instance, as RankingMetric does
                order = torch.argsort(logits, dim=1, descending=True, stable=True)[:, :cfg["k"]]
                hit = order.eq(target.unsqueeze(1))
                if name == "recall":
                    out[name] = hit.any(1).float().mean()
                else:
                    pos = hit.float().argmax(1)
                    out[name] = (hit.any(1).float() / torch.log2(pos.float() + 2.0)).mean()
        return float(out["recall"])

    return step


def time_cpu(step, units, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return units / med, med, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(5, min(args.steps, 30))
    warmup = max(2, min(args.warmup, 5))
    value, t_step, cores = time_cpu(oracle_train_setup(C2, C2["B"]), C2["B"], steps, warmup)
    sample = (f"the full C2 batch ({C2['B']} sequences, 4 distinct cloze batches rotating) per step, median of {steps} steps after {warmup} warm-up "
              f"({t_step:.2f} s/step): oracle/asme_oracle.py port of the reference's PyTorch path, dropout 0.2, fwd + CE + bwd + Adam")
    ev = None
    if not args.no_eval:
        ev_v, ev_t, _ = time_cpu(oracle_eval_setup(C5, C5["B_cpu"]), C5["B_cpu"], 5, 2)
        ev = {"metric": "eval_users_per_sec", "value": ev_v, "unit": "users/s", "ms_per_step": ev_t * 1e3, "users_per_step": C5["B_cpu"],
              "note": "B=64 users per step (SURVEY 8d): encoder + MASK-row select + projection of the selected rows onto 1M items + one "
                      "full argsort per metric"}
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": workload_config(max(1, args.gpus)),
           "implementation": {"precision": "fp32 (torch CPU ops)", "launch": "oracle/asme_oracle.py on the host cores", "distinct_batches": 4},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
           "timed_steps": steps, "eval": ev}
    emit(out)


# ----------------------------------------------------------------------------------------------------------------
# roofline model: algorithmic bytes / flops per launch of each C-ABI entry point, from the shapes noted by ops.py
# ----------------------------------------------------------------------------------------------------------------
def algorithmic_cost(name, note):
    """returns (bytes, flops) per call; shapes come from ``note``.  DESIGN.md section 4 states the per-unit figures."""
    f = dict(kv.split("=") for kv in note.split(",") if "=" in kv) if note else {}
    g = lambda k: int(f.get(k, 0))
    if name == "asme_b200_gemm":
        M, N, K = g("M"), g("N"), g("K")
        extra = g("res") + g("pre") + g("aux")
        return 4 * (M * K + N * K + M * N * (1 + extra)), 2 * M * N * K
    if name == "asme_b200_gemm_wgrad":
        M, N, K = g("M"), g("N"), g("K")
        return 4 * (M * N + M * K + N * K), 2 * M * N * K
    if name in ("asme_b200_tc_gemm", "asme_b200_tc_gemm_ln"):
        M, N, K = g("M"), g("N"), g("K")
        b = 2 * (M * K + N * K) + M * N * (4 * g("f32") + 2 * g("bf16") + 2 * g("pre") + 2 * g("aux") + 4 * g("res"))
        return b, 2 * M * N * K
    if name == "asme_b200_tc_wgrad":
        M, N, K = g("M"), g("N"), g("K")
        return 2 * (M * N + M * K) + 4 * N * K, 2 * M * N * K
    if name == "asme_b200_tc_attn_fwd":
        T, H, S = g("T"), g("H"), g("S")
        return 2 * 4 * T * H, 4 * T * S * H
    if name == "asme_b200_tc_attn_bwd":
        T, H, S = g("T"), g("H"), g("S")
        return 2 * 9 * T * H, 16 * T * S * H
    if name == "asme_b200_cast_bf16":
        n = g("rows") * g("ld")
        return 6 * n, n
    if name == "asme_b200_dropout_cast":
        n = g("n")
        return (4 + 2 + 4 * g("f32")) * n, n
    if name in ("asme_b200_tc_score_topk", "asme_b200_tc_score_candidates"):
        # flops on the MODEL's hidden size: the bias-folding columns (Kp = H + 16) are this implementation's overhead, not algorithmic work
        R, V, Kp, k = g("R"), g("V"), g("H"), g("k")
        H = Kp - 16 if Kp % 64 == 16 else Kp
        return 2 * (R * H + V * H) + 4 * V + R * (k * 8 + 8), 2 * R * V * H
    if name == "asme_b200_tc_score_ce_bwd":
        R, V, H = g("R"), g("V"), g("H")
        return 2 * 2 * (R * H + V * H) + 4 * (R * H + 2 * V * H + 3 * V), 12 * R * V * H
    if name == "asme_b200_tc_score_ce_partial":
        R, V, H = g("R"), g("V"), g("H")
        return 2 * (R * H + V * H) + 4 * V + 12 * R, 2 * R * V * H
    if name == "asme_b200_attn_fwd":
        T, H, S = g("T"), g("H"), g("S")
        return 4 * 4 * T * H, 4 * T * S * H
    if name == "asme_b200_attn_bwd":
        T, H, S = g("T"), g("H"), g("S")
        return 4 * 8 * T * H, 10 * T * S * H
    if name == "asme_b200_layernorm_fwd":
        M, H = g("M"), g("H")
        return 4 * 2 * M * H, 8 * M * H
    if name == "asme_b200_layernorm_fwd_bf16":       # reads fp32 rows, writes bf16 rows (and an fp32 copy when asked)
        M, H = g("M"), g("H")
        return (4 + 2 + 4 * g("f32")) * M * H, 8 * M * H
    if name == "asme_b200_tc_block_tail_fused":      # ctx (bf16) + the block's input rows (fp32) in, fp32 rows and / or LayerNorm'ed bf16 rows out;
        M, H, FF = g("M"), g("H"), g("FF")           # x2, y and the (M, FF) intermediate are not algorithmic traffic
        return M * H * (2 + 4 + 4 * g("f32") + 2 * g("ln")) + 2 * (H * H + 2 * H * FF), 2 * M * H * H + 4 * M * H * FF
    if name == "asme_b200_tc_ffn_fused":             # y (bf16) + residual (fp32) in, fp32 rows and / or LayerNorm'ed bf16 rows out; the
        M, H, FF = g("M"), g("H"), g("FF")           # (M, FF) intermediate is not algorithmic traffic; both weight matrices once
        return M * H * (2 + 4 + 4 * g("f32") + 2 * g("ln")) + 2 * 2 * H * FF, 4 * M * H * FF
    if name == "asme_b200_layernorm_bwd":
        M, H = g("M"), g("H")
        return 4 * (3 + g("res")) * M * H, 16 * M * H
    if name == "asme_b200_layernorm_bwd_drop":      # reads dy, x (+ residual), writes dx fp32 and its dropped bf16 copy
        M, H = g("M"), g("H")
        return (4 * (3 + g("res")) + 2) * M * H, 20 * M * H
    if name == "asme_b200_attn_row_fwd":     # K and V rows of every token once (bf16), one query / context row per sequence
        T, H, S = g("T"), g("H"), g("S")
        return 2 * T * H * 2 + T + 2 * (T // max(S, 1)) * H * 2, 4 * T * H
    if name == "asme_b200_embed_fwd":
        T, H, nt = g("T"), g("H"), g("tables")
        return nt * T * H * 4 + g("ids") * T * 8 + T * H * 4 + g("next") * T * H * 2, 8 * T * H
    if name == "asme_b200_embed_bwd":
        T, H, nt = g("T"), g("H"), g("tables")
        return (nt + 2) * T * H * 4 + g("ids") * T * 8, 16 * T * H
    if name == "asme_b200_score_ce_partial":
        R, V, H = g("R"), g("V"), g("H")
        return 4 * (R * H + V * H + V + 3 * R), 2 * R * V * H
    if name == "asme_b200_score_ce_bwd":
        R, V, H = g("R"), g("V"), g("H")
        return 4 * (2 * R * H + 2 * V * H + 2 * V), 8 * R * V * H
    if name in ("asme_b200_score_topk_rank", "asme_b200_score_topk_flagged"):
        R, V, H, k = g("R"), g("V"), g("H"), g("k")
        if g("flagged"):
            return R * (k * 8 + 8), 0          # nothing flagged: the launch only reads the flags
        return 4 * (R * H + V * H + V) + R * (k * 8 + 8), 2 * R * V * H
    if name == "asme_b200_topk_rescore":     # candidate rows of the fp32 table + the hidden row, per user
        R, H, KC = g("R"), g("H"), g("KC")
        return R * (KC * (H * 4 + 8) + H * 4), 2 * R * KC * H
    if name == "asme_b200_embgrad_sorted_reduce":
        T, H = g("T"), g("H")
        return T * H * 4 * 2 + T * 8, T * H
    if name == "asme_b200_embgrad_sort":          # (id, token) keys: read ids, radix sort passes over 8-byte pairs
        T = g("T")
        return T * 8 + 4 * T * 8 * 2, T
    if name == "asme_b200_embgrad_reduce_sorted":  # gradient rows read once, one row written per distinct id (<= T)
        T, H = g("T"), g("H")
        return T * H * 4 * 2 + T * 8, T * H
    if name in ("asme_b200_gather_rows", "asme_b200_scatter_rows"):
        R, H = g("R"), g("H")
        return R * (2 * H * 4 + 8), 0
    if name == "asme_b200_select_rows":
        T = g("T")
        return T * 8 * 2 + T * 16, T
    if name == "asme_b200_fill":
        return 4 * g("n"), 0
    if name == "asme_b200_ce_loss_from_partials":
        return 4 * 4 * g("R"), 4 * g("R")
    if name in ("asme_b200_posgrad_reduce", "asme_b200_posgrad_reduce_strided"):
        B, S, H = g("B"), g("S"), g("H")
        return 4 * (B * S * H + S * H), B * S * H
    if name == "asme_b200_colsum_accumulate":
        return 4 * g("M") * g("N"), g("M") * g("N")
    if name in ("asme_b200_dropout", "asme_b200_gelu_bwd", "asme_b200_binary"):
        n = g("n")
        return 4 * 2 * n, n
    if name in ("asme_b200_adam_step", "asme_b200_adam_step_dev"):
        n = g("n")
        return 4 * 7 * n, 12 * n
    return 0, 0


NCU_NAMES = {"tc_attn_bwd": "attn_tc_bwd1_kernel", "tc_attn_fwd": "attn_tc_fwd_kernel", "tc_gemm": "tc_gemm_persist_kernel",
             "tc_wgrad": "tc_wgrad_kernel", "tc_score_topk": "score_tc_kernel", "tc_score_candidates": "score_tc_kernel",
             "tc_score_ce_bwd": "ce_bwd_tc_kernel", "tc_score_ce_partial": "score_tc_kernel"}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel behind a C-ABI entry point, from the committed
    `ncu --set full` summary of this same step (profiles/*_ncu_full.json); None when there is no capture of it"""
    import glob
    want = NCU_NAMES.get(kernel)
    if want is None:
        return None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_full.json")), reverse=True):
        try:
            recs = json.load(open(path))
        except Exception:
            continue
        for r in recs:
            if want in r.get("kernel", ""):
                def mb(x):
                    v, u = x.split()[:2]
                    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                try:
                    return {"bytes_per_launch": mb(r["dram__bytes_read.sum"]) + mb(r["dram__bytes_write.sum"]), "source": os.path.basename(path)}
                except Exception:
                    continue          # this capture has no DRAM counters for the kernel: try the next record / older capture
    return None


def summarise_kernels(records, steps, pk):
    """records: (name, note, ms). Returns per-entry-point table (sorted by share) and the kernel time per step."""
    groups = {}
    for name, note, ms in records:
        gsum = groups.setdefault((name, note), [0.0, 0])
        gsum[0] += ms
        gsum[1] += 1
    total = sum(v[0] for v in groups.values())
    table = []
    for (name, note), (ms, n) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        b, fl = algorithmic_cost(name, note)
        avg_ms = ms / n
        t_hbm = b / (pk["hbm_gbs"] * 1e9)
        t_tc = fl / (pk["bf16_tflops"] * 1e12)
        bound = "hbm" if t_hbm >= t_tc else "tensor"
        achieved = (b / 1e9) / (avg_ms / 1e3) if bound == "hbm" else (fl / 1e12) / (avg_ms / 1e3)
        peak = pk["hbm_gbs"] if bound == "hbm" else pk["bf16_tflops"]
        table.append(dict(kernel=name.replace("asme_b200_", ""), shape=note, launches_per_step=n / steps, avg_ms=avg_ms,
                          share=ms / total if total else 0.0, bound=bound, achieved=achieved, peak=peak,
                          unit="GB/s" if bound == "hbm" else "TFLOP/s", frac=achieved / peak if peak else None,
                          algorithmic_bytes=b, algorithmic_flops=fl))
    return table, total / steps


def roofline_of(row, pk):
    traffic = ncu_traffic(row["kernel"]) or {}
    return {"kernel": row["kernel"], "shape": row["shape"], "bound": row["bound"], "achieved": row["achieved"], "peak": row["peak"],
            "unit": row["unit"], "frac": row["frac"], "traffic": traffic.get("bytes_per_launch"), "traffic_source": traffic.get("source"),
            "avg_launch_ms": row["avg_ms"], "share_of_step": row["share"], "peak_source": pk["source"],
            "algorithmic_bytes_per_launch": row["algorithmic_bytes"], "algorithmic_flops_per_launch": row["algorithmic_flops"]}


# ----------------------------------------------------------------------------------------------------------------
_RESULT_FD = None


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def log_table(title, table):
    rows = [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in row.items()} for row in table]
    print(f"[bench kernels] {title}: " + json.dumps(rows), file=sys.stderr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-eval", action="store_true", help="skip the C5 full-catalog evaluation half")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baselines")
    ap.add_argument("--no-graph", action="store_true", help="issue the steps launch by launch instead of replaying their CUDA graphs")
    ap.add_argument("--profile-region", action="store_true",
                    help="profiling aid (ncu --profile-from-start off): after warm-up bracket ONE launch-by-launch training step and ONE "
                         "evaluation step in cudaProfilerStart/Stop, print no bench line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 on their own (NCCL prints its version
    # banner there) are sent to stderr for the whole run, the result line goes to the saved descriptor
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    from asme_b200 import _lib, input_pipeline
    from asme_b200.metrics import build_metrics
    from asme_b200.models import BERT4RecModel
    from asme_b200.modules import MaskedTrainingModule
    _lib.load()
    pk = peaks()
    cfg = C2
    torch.manual_seed(0)
    model = BERT4RecModel(cfg["H"], cfg["heads"], cfg["L"], cfg["V"], cfg["S"], cfg["dropout"])
    module = MaskedTrainingModule(model, metrics=build_metrics({"recall": [1, 5, 10], "ndcg": [1, 5, 10]}), learning_rate=1e-3,
                                  num_warmup_steps=100).to(device)
    module.train()
    (optimizer,), (sched,) = module.configure_optimizers()
    scheduler = sched["scheduler"]

    # 64 distinct batches per rank (data parallel: every rank owns its own sequences).  Host side: the unmasked item sequences in
    # pinned memory, as the session store delivers them; device side: cloze-masked by the GPU input pipeline (csrc/pipeline.cu) --
    # every batch draws its own number of masked positions
    gen = torch.Generator().manual_seed(1234 + 1 + 1000 * rank)
    host_items = [make_sessions(gen, cfg["B"], cfg["S"], cfg["V"], cfg["min_len"]).pin_memory() for _ in range(N_BATCHES)]

    def cloze(items_dev, j):
        return input_pipeline.cloze_mask({"item": items_dev}, {"item": cfg["V"]}, cfg["mask_prob"], cfg["only_last_prob"],
                                         seed=(rank << 20) + j)

    dev_batches = [cloze(h.to(device), j) for j, h in enumerate(host_items)]
    target_counts = sorted({int((b["item.target"] != 0).sum()) for b in dev_batches})

    def allreduce_grads():
        if world > 1:
            g = model._arena.ensure_grad()
            dist.all_reduce(g)                     # ONE NCCL call over the flat gradient arena
            g.div_(world)

    def train_step(batch, i):
        optimizer.zero_grad()
        out = module.training_step(batch, i)
        out["loss"].backward()
        allreduce_grads()
        optimizer.step()
        scheduler.step()
        return out["loss"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.profile_region:
        for i in range(3):
            train_step(dev_batches[i], i)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        train_step(dev_batches[3], 3)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        bench_eval_c5(args, device, pk, world=world, rank=rank, profile=True)
        emit({"profile_region": "1 training step (C2) + 1 evaluation step (C5), launch by launch"})
        return

    # The step is ~90 launches; issued from Python the host is the bottleneck, so the product path replays the step's CUDA graph
    # (asme_b200.graphs.GraphedTrainStep).  ONE graph per batch shape: the positions that carry a target are selected on the device.
    graphed = None
    if not args.no_graph:
        from asme_b200.graphs import GraphedTrainStep
        try:
            graphed = GraphedTrainStep(module, optimizer, scheduler, grad_hook=allreduce_grads if world > 1 else None)
            graphed(dev_batches[0])
            torch.cuda.synchronize()
        except Exception as e:      # e.g. a collective that cannot be captured: fall back to launch-by-launch
            print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            graphed = None
            module.model._step_state = None
            optimizer.step_state = None

    def run_step(batch, i):
        return graphed(batch) if graphed is not None else train_step(batch, i)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        t_wall1 = time.time()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, (t_wall0, t_wall1)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()

    # ---- (1) device-resident throughput: a different batch every step --------------------------------------------------------------
    for i in range(args.warmup):
        run_step(dev_batches[i % N_BATCHES], i)
    ms, win1 = timed(lambda i: run_step(dev_batches[(args.warmup + i) % N_BATCHES], i), args.steps)
    value = args.steps * cfg["B"] * world / (ms / 1e3)

    # ---- (2) end to end through the public API with HOST buffers ---------------------------------------------------------------------
    # The loop a trainer with a prefetching loader runs: the H2D copy of step i+1's item sequences is issued on a copy stream while step
    # i runs (two staging buffers; an event hands each back once the cloze kernels have read it), and the loss of step i is copied to
    # pinned memory behind the step and READ after step i+1 has been queued, so neither direction stalls the device.  Every step
    # still copies its own inputs from pinned host memory and its own loss back; the last loss is read inside the timed region.
    h2d = host_items[0].numel() * host_items[0].element_size()
    staged = [torch.empty_like(dev_batches[0]["item"]) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=device)
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event() for _ in range(2)]
    e2e_losses = []

    def make_e2e(n):
        def prefetch(i):
            b = i & 1
            copy_stream.wait_event(in_free[b])                                # its previous reader (step i-2's cloze kernels) is done
            with torch.cuda.stream(copy_stream):
                staged[b].copy_(host_items[(args.warmup + i) % N_BATCHES], non_blocking=True)    # H2D of step i's item sequences
                in_ready[b].record(copy_stream)

        def e2e_step(i):
            b = i & 1
            if i == 0:
                for ev in in_free:
                    ev.record()
                prefetch(0)
            torch.cuda.current_stream().wait_event(in_ready[b])
            batch = cloze(staged[b], N_BATCHES + i)                           # input pipeline on the GPU: fresh masks every step
            in_free[b].record()
            if i + 1 < n:
                prefetch(i + 1)                                               # overlaps this step
            loss = run_step(batch, i)                                         # the step
            if not torch.is_tensor(loss):
                loss = torch.tensor(float(loss), device=device)
            loss_host[b].copy_(loss.detach().reshape(()).float(), non_blocking=True)  # D2H of its loss, behind it on the same stream
            loss_done[b].record()
            if i > 0:                                                         # read the PREVIOUS step's loss: it has long arrived
                loss_done[b ^ 1].synchronize()
                e2e_losses.append(float(loss_host[b ^ 1]))
            if i == n - 1:
                loss_done[b].synchronize()
                e2e_losses.append(float(loss_host[b]))
        return e2e_step

    warm = make_e2e(3)
    for i in range(3):
        warm(i)
    e2e_losses.clear()
    ms_e2e, win2 = timed(make_e2e(args.steps), args.steps)
    assert len(e2e_losses) == args.steps and all(math.isfinite(x) for x in e2e_losses), "end-to-end loop lost a loss value"
    e2e_value = args.steps * cfg["B"] * world / (ms_e2e / 1e3)
    graphs_captured = len(graphed.graphs) if graphed is not None else 0

    # ---- (3) per-kernel CUDA-event timing of the same step (roofline of the dominant kernel) -----------------------------------------
    prof_steps = min(5, args.steps)
    if graphed is not None:          # the per-kernel pass (and the launch count) runs the same step launch by launch
        module.model._step_state = None
        optimizer.step_state = None
    _lib.timing = []
    _lib.timing_spacer_cycles = 120000          # ~60 us of spin before every timed call: host issue gaps stay out of the kernel times
    barrier()
    l0 = _lib.kernel_launches()
    for i in range(prof_steps):
        train_step(dev_batches[i % N_BATCHES], i)
    launches = (_lib.kernel_launches() - l0) * args.steps // prof_steps   # kernels of this library per step x timed steps (graph replays launch the same nodes)
    torch.cuda.synchronize()
    records = [(n, note, a.elapsed_time(b)) for (n, note, a, b) in _lib.timing]
    _lib.timing = None
    _lib.timing_spacer_cycles = 0
    table, kernel_ms_per_step = summarise_kernels(records, prof_steps, pk)
    unaccounted = sum(r["share"] for r in table if r["algorithmic_bytes"] == 0 and r["algorithmic_flops"] == 0)
    # the timed regions last tens of milliseconds -- shorter than nvidia-smi's sampling period -- so the same step keeps
    # running (untimed) until enough clock samples were taken under exactly this load
    probe = None
    if rank == 0 and clocks.proc is not None:
        t_probe0 = time.time()
        i = 0
        while clocks.count_in([win1, win2, (t_probe0, time.time())]) < 8 and time.time() - t_probe0 < 4.0:
            optimizer.zero_grad()                      # forward + backward only: no collective, weights stay in sync across ranks
            module.training_step(dev_batches[i % N_BATCHES], i)["loss"].backward()
            i += 1
            if i % 20 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        probe = (t_probe0, time.time())
    if world > 1:
        dist.barrier()
    clocks.stop()

    # ---- (4) the evaluation half: C5 full-catalog scoring + exact top-k + Recall/NDCG ------------------------------------------------
    eval_info = None
    if not args.no_eval:
        del dev_batches
        eval_info = bench_eval_c5(args, device, pk, world=world, rank=rank, graph=not args.no_graph)

    def finish():
        # CUDA graphs that captured NCCL collectives keep communicator resources alive and ncclCommDestroy can then block for ever:
        # flush what was printed and leave without tearing the process group down.
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        finish()
        return

    roofline = roofline_of(table[0], pk) if table else None
    log_table("C2 training step", table[:40])
    cpu = None
    if not args.no_cpu and world == 1:
        v_cpu, t_cpu, cores = time_cpu(oracle_train_setup(cfg, cfg["B"]), cfg["B"], steps=5, warmup=2)
        cpu = {"value": v_cpu, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"the full C2 batch ({cfg['B']} sequences) per step, median of 5 steps after 2 warm-up ({t_cpu:.2f} s/step); "
                         f"oracle/asme_oracle.py port of the reference path, dropout 0.2, fwd+CE+bwd+Adam"}
    launch = ("CUDA graph replay of the whole step: ONE graph for every batch of the shape (target rows selected on the device)"
              if graphed is not None else "launch by launch")
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16" if model.precision == "bf16" else "f32", "data": "synthetic",
           "config": workload_config(world),
           "implementation": {"precision": f"{model.precision}: tcgen05 GEMMs/attention/scoring with bf16 operands, fp32 accumulation, residual "
                                           f"stream, LayerNorm, loss and Adam state", "launch": launch, "distinct_batches": N_BATCHES},
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                   "path": "pinned host item sequences -> H2D (copy stream, double-buffered: the copy of step i+1 overlaps step i) -> GPU cloze masking -> MaskedTrainingModule step (graph replay) -> loss D2H to pinned memory, read one step later (the last one inside the timed region)"},
           "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
           "graphs_captured": graphs_captured, "distinct_target_counts": len(target_counts),
           "target_count_range": [target_counts[0], target_counts[-1]],
           "clocks": dict(clocks.summary([win1, win2] + ([probe] if probe else [])),
                          window="timed regions + an untimed continuation of the same step until 8 samples (nvidia-smi -lms 50)"),
           "roofline": roofline, "cpu_baseline": cpu, "kernel_ms_per_step": kernel_ms_per_step,
           "kernel_time_without_cost_model": unaccounted,
           "kernel_timing": "CUDA events around every C-ABI call of one launch-by-launch step on the launching stream, a ~60 us spin kernel "
                            "queued ahead of each call so that the host's issue gap is not charged to the kernel",
           "kernels_top5": [{k: (round(r[k], 4) if isinstance(r[k], float) else r[k]) for k in ("kernel", "avg_ms", "share", "bound", "frac")}
                            for r in table[:5]],
           "eval": eval_info}
    if eval_info:        # the evaluation half once more as the LAST key of the line, compact (whoever keeps only the tail of the line sees it)
        ev = eval_info
        out["eval_summary"] = {"metric": ev["metric"], "value": ev["value"], "unit": ev["unit"], "ms_per_step": ev["ms_per_step"],
                               "n_gpus": ev["n_gpus"], "users_per_step_per_gpu": ev["users_per_step_per_gpu"],
                               "e2e_value": ev["e2e"]["value"], "e2e_h2d_bytes_per_step": ev["e2e"]["h2d_bytes_per_step"],
                               "recall@10": ev["recall@10"], "NDCG@10": ev["NDCG@10"], "exact_topk": ev["exact_topk"],
                               "scoring_frac_of_tensor_peak": None if not ev.get("roofline") else ev["roofline"]["frac"],
                               "cpu_baseline_value": None if not ev.get("cpu_baseline") else ev["cpu_baseline"]["value"],
                               "checks": ev.get("checks")}
    emit(out)
    finish()


def bench_eval_c5(args, device, pk, world=1, rank=0, profile=False, graph=True):
    """the evaluation half of the metric: full-catalog scoring + EXACT top-10 + Recall/NDCG@10 on a 1M-item catalog (C5), through
    ``MaskedTrainingModule.validation_step`` / ``validation_step_end``.  With N ranks the catalog is vocab-sharded
    (asme_b200.sharded): every rank encodes its own 1024 users and scores all N*1024 users against its V/N slice (weak scaling)."""
    import torch.distributed as dist
    from asme_b200 import _lib
    from asme_b200.metrics import build_metrics
    from asme_b200.models import BERT4RecModel
    from asme_b200.modules import MaskedTrainingModule
    cfg = C5
    torch.manual_seed(0)
    model = BERT4RecModel(cfg["H"], cfg["heads"], cfg["L"], cfg["V"], cfg["S"], 0.0)
    module = MaskedTrainingModule(model, metrics=build_metrics({"recall": [10], "ndcg": [10]})).to(device).eval()
    module.eval_loss = False
    module.shard_catalog = world > 1
    module.eval_graph = bool(graph and not profile)
    gen = torch.Generator().manual_seed(1234 + 4 + 1000 * rank)
    B, S, V, k = cfg["B"], cfg["S"], cfg["V"], cfg["k"]
    seq = make_sessions(gen, B, S - 1, V, 20)
    seq = torch.cat([seq, torch.zeros(B, 1, dtype=seq.dtype)], dim=1)
    seq[torch.arange(B), seq.ne(0).sum(1)] = 1                       # one MASK appended per user
    free = torch.randint(3, V, (B,), generator=gen)
    seq_d = seq.to(device)
    # targets: for every second user an item of the user's OWN top 20 (so that Recall@10 is not trivially zero on random weights)
    with torch.no_grad():
        first = model.evaluate_rank(seq_d, seq_d.ne(0), {}, free.to(device), k=20)
    col = torch.randint(0, 20, (B,), generator=gen).to(device)
    planted = first["topk_idx"].gather(1, col.unsqueeze(1)).squeeze(1).to(torch.int64)
    target_d = torch.where(torch.arange(B, device=device) % 2 == 0, planted, free.to(device))
    host = {"item": seq.pin_memory(), "item.target": target_d.cpu().pin_memory()}
    dev = {"item": seq_d, "item.target": target_d}
    staged = {k_: torch.empty_like(v) for k_, v in dev.items()}

    def step(batch):
        with torch.no_grad():
            return module.validation_step_end(module.validation_step(batch, 0))

    def e2e_step():
        for k_ in staged:
            staged[k_].copy_(host[k_], non_blocking=True)               # H2D of this step's users (pinned)
        values = step(staged)
        return float(values["recall@10"]), float(values["NDCG@10"])     # D2H of the step's metric values

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    steps, warmup = max(5, min(args.steps, 20)), max(3, min(args.warmup, 5))
    for _ in range(warmup):
        step(dev)
    barrier()
    if profile:
        torch.cuda.profiler.start()
        step(dev)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(dev)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / steps)
    for _ in range(3):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(steps):
        recall_step, ndcg_step = e2e_step()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1) / steps)
    module.metrics.sync() if world > 1 else None
    res = module.validation_epoch_end(None)

    # ---- checks, outside the timed regions -----------------------------------------------------------------------------------------
    checks = {}
    with torch.no_grad():
        mine = model.evaluate_rank(seq_d, seq_d.ne(0), {}, target_d, k=k, full_rank=False)
        if world > 1:       # the sharded result of this rank's users equals its own unsharded evaluation
            sh = model.evaluate_rank_sharded(seq_d, seq_d.ne(0), {}, target_d, k=k, full_rank=False)
            checks["sharded_equals_unsharded"] = bool(torch.equal(sh["topk_idx"], mine["topk_idx"]) and torch.equal(sh["rank"], mine["rank"])
                                                      and torch.equal(sh["topk_val"], mine["topk_val"]))
        # dense cross-check on a row subset: float64 logits of the fp32 operands -> stable sort -> Recall@10
        from asme_b200.models import mask_position_rows
        n_chk = 64
        m_rows, _ = model.modify(model.encode_rows(seq_d, seq_d.ne(0), {}, mask_position_rows(seq_d, 1), one_per_sequence=True))
        w32, b32 = model.projection_operands()
        logits = m_rows[:n_chk].double() @ w32.double().t() + b32.double()
        order = torch.argsort(logits, dim=1, descending=True, stable=True)[:, :k]
        checks["dense_rows"] = n_chk
        checks["topk_equals_dense_fp64"] = bool(torch.equal(order.to(torch.int32), mine["topk_idx"][:n_chk]))
        hit = order.eq(target_d[:n_chk].unsqueeze(1)).any(1)
        checks["recall_equals_dense"] = bool(torch.equal(hit, mine["rank"][:n_chk] <= k))
        checks["uncertified_rows"] = int(mine["n_uncertified"]) if "n_uncertified" in mine else None
    # ---- per-kernel pass of one launch-by-launch step -----------------------------------------------------------------------------
    module.eval_graph = False
    _lib.timing = []
    _lib.timing_spacer_cycles = 120000          # ~60 us of spin before every timed call: host issue gaps stay out of the kernel times
    step(dev)
    torch.cuda.synchronize()
    rec = [(n, note, a.elapsed_time(b)) for (n, note, a, b) in _lib.timing]
    _lib.timing = None
    _lib.timing_spacer_cycles = 0
    table, kernel_ms = summarise_kernels(rec, 1, pk)
    score = next((r for r in table if r["kernel"] in ("tc_score_candidates", "tc_score_topk", "score_topk_rank")), None)
    # ---- the other batch size of SURVEY 8d (4096 users per step): a second data point, the headline stays at 1024 ------------------
    big = None
    if world == 1 and not args.no_eval:
        try:
            B4 = 4 * B
            seq4 = torch.cat([seq] + [torch.roll(seq, shifts=r, dims=0) for r in (1, 2, 3)], dim=0)
            tgt4 = torch.cat([target_d] * 4, dim=0)
            batch4 = {"item": seq4.to(device), "item.target": tgt4}
            module.eval_graph = bool(graph)
            for _ in range(3):
                step(batch4)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                step(batch4)
            e1.record()
            torch.cuda.synchronize()
            ms4 = e0.elapsed_time(e1) / 10
            _lib.timing = []
            _lib.timing_spacer_cycles = 120000
            module.eval_graph = False
            step(batch4)
            torch.cuda.synchronize()
            rec4 = [(n, note, a.elapsed_time(b)) for (n, note, a, b) in _lib.timing]
            _lib.timing = None
            _lib.timing_spacer_cycles = 0
            t4, _ = summarise_kernels(rec4, 1, pk)
            s4 = next((r for r in t4 if r["kernel"] == "tc_score_candidates"), None)
            big = {"users_per_step": B4, "value": B4 / (ms4 / 1e3), "unit": "users/s", "ms_per_step": ms4,
                   "scoring_call_ms": None if s4 is None else round(s4["avg_ms"], 4), "scoring_frac_of_tensor_peak": None if s4 is None else round(s4["frac"], 4)}
            module.metrics.reset() if hasattr(module.metrics, "reset") else None
        except Exception as ex:          # the extra data point must never take the bench line down
            big = {"error": repr(ex)[:200]}
            _lib.timing = None
            _lib.timing_spacer_cycles = 0
    if rank == 0:
        log_table("C5 evaluation step", table[:30])
    if rank != 0:
        return None
    cpu = None
    if not args.no_cpu and world == 1:
        v_cpu, t_cpu, cores = time_cpu(oracle_eval_setup(cfg, cfg["B_cpu"]), cfg["B_cpu"], steps=5, warmup=2)
        cpu = {"value": v_cpu, "unit": "users/s", "cores": cores, "kind": "port",
               "sample": f"{cfg['B_cpu']} users per step (SURVEY 8d), median of 5 after 2 warm-up ({t_cpu:.2f} s/step): encoder, MASK-row select, "
                         f"projection of the selected rows onto the 1M catalog, one full argsort per metric"}
    roof = None
    if score is not None:
        roof = roofline_of(score, pk)
        roof["note"] = ("whole candidates call (threshold pass + merge + main sweep + merge); flops = 2 * users * items * 128 (the model's hidden "
                        "size; the kernel contracts Kp = 144 columns: bias folded into the table)")
    return {"metric": "eval_users_per_sec", "value": world * B / (ms / 1e3), "unit": "users/s", "ms_per_step": ms, "n_gpus": world, "steps": steps,
            "scaling": "weak", "users_per_step_per_gpu": B, "items": V, "hidden": cfg["H"], "seq_len": S, "k": k,
            "launch": "graph replay of the model part" if module.__dict__.get("_eval_graphs") else "launch by launch",
            "sharding": "single GPU" if world == 1 else f"catalog rows over {world} ranks: 1 all-gather (hidden rows+targets), 1 all-to-all (lists of own users)",
            "exact_topk": bool(model.exact_topk), "recall@10": float(res["recall@10"]), "NDCG@10": float(res["NDCG@10"]),
            "e2e": {"value": world * B / (ms_e2e / 1e3), "unit": "users/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values()), "d2h_bytes_per_step": 8,
                    "path": "pinned host (item, item.target) -> H2D -> MaskedTrainingModule.validation_step + step_end -> metric values D2H"},
            "roofline": roof, "cpu_baseline": cpu, "checks": checks, "kernel_ms_per_step": kernel_ms, "batch_4096": big,
            "kernels_top5": [{k_: (round(r[k_], 4) if isinstance(r[k_], float) else r[k_]) for k_ in ("kernel", "avg_ms", "share", "bound", "frac")}
                             for r in table[:5]]}


if __name__ == "__main__":
    main()
