"""Host-side throughput of the vectorised collate (asme_b200.formats.TokenisedSessions.batch) at the C2 shape, next to a literal
per-sample restatement of what the reference does per batch (slice, target, truncate, pad with python lists, torch.as_tensor,
stack) on ALREADY tokenised sessions -- i.e. without the reference's per-sample file open / csv parse / string tokenisation."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")]
import numpy as np
import torch
from asme_b200.formats import TokenisedSessions

rng = np.random.default_rng(0)
n_sessions, B, S = 6040, 256, 200
lengths = np.clip(rng.lognormal(4.6, 0.9, n_sessions).astype(np.int64), 20, 2314)          # ML-1M-like session lengths
offsets = np.concatenate([[0], np.cumsum(lengths)])
store = TokenisedSessions(rng.integers(3, 3709, offsets[-1]).astype(np.int32), offsets)
sid = rng.integers(0, n_sessions, (200, B))
pos = np.stack([rng.integers(1, lengths[s]) for s in sid.reshape(-1)]).reshape(sid.shape)


def vectorised(i):
    return store.batch(sid[i], pos[i], max_seq_length=S)


def per_sample(i):
    items, targets = [], []
    for s, p in zip(sid[i], pos[i]):
        seq = store.sequence(int(s)).tolist()[:p + 1]
        targets.append(seq[-1])
        x = seq[:-1][-S:]
        items.append(torch.as_tensor(x + [0] * (S - len(x))))
    return {"item": torch.stack(items), "item.target": torch.as_tensor(targets)}


for name, fn, n in (("vectorised", vectorised, 200), ("per_sample", per_sample, 20)):
    fn(0)
    t0 = time.perf_counter()
    for i in range(n):
        out = fn(i)
    dt = (time.perf_counter() - t0) / n
    print(json.dumps(dict(path=name, ms_per_batch=round(dt * 1e3, 3), sequences_per_s=round(B / dt), tokens_per_s=round(B * S / dt))))
a, b = vectorised(3), per_sample(3)
assert torch.equal(a["item"], b["item"]) and torch.equal(a["item.target"], b["item.target"])
