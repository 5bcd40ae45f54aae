"""Golden output of the `predict` command's evaluators + CSV writers, produced by the UNMODIFIED reference classes
(asme/core/evaluation/evaluation.py, asme/core/writer/prediction/batch_prediction_writer.py) on seeded dense logits:

    python tests/golden/make_predict_golden.py      ->  tests/golden/predict_small.json

The test feeds the same logits (dense, and as top-n list + log-sum-exp) to asme_b200.evaluation and compares the CSV text.
This script is the only code that touches /root/reference."""
import io
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shims"))
import ref_shims  # noqa: E402

ref_shims.install()

from asme.core.evaluation.evaluation import (ExtractRecommendationEvaluator, ExtractSampleIdEvaluator, ExtractScoresEvaluator,  # noqa: E402
                                             LogInputEvaluator, TrueTargetEvaluator)
from asme.core.tokenization.tokenizer import Tokenizer  # noqa: E402
from asme.core.tokenization.vocabulary import VocabularyBuilder  # noqa: E402
from asme.core.writer.prediction.batch_prediction_writer import CSVMultiLineWriter, CSVSingleLineWriter  # noqa: E402

V, N, n = 30, 5, 4
vb = VocabularyBuilder()
tokens = ["<PAD>", "<MASK>", "<UNK>"] + [f"Item {i}" for i in range(3, V)]
for t in tokens:
    vb.add_token(t)
tokenizer = Tokenizer(vb.build(), pad_token="<PAD>", mask_token="<MASK>", unk_token="<UNK>")
gen = torch.Generator().manual_seed(11)
logits = torch.randn(N, V, generator=gen) * 3          # well separated: no ties, no near-ties
batch = {"item": torch.tensor([[5, 7, 1, 0, 0], [9, 1, 0, 0, 0], [3, 4, 6, 8, 1], [8, 1, 0, 0, 0], [11, 12, 13, 1, 0]]),
         "item.target": torch.tensor([4, 8, 15, 16, 23]), "sample_ids": torch.arange(100, 100 + N), "pos": torch.tensor([2, 1, 4, 1, 3])}


def run(writer_cls, selected=None):
    evaluators = [ExtractSampleIdEvaluator(), LogInputEvaluator(tokenizer), TrueTargetEvaluator(tokenizer),
                  ExtractRecommendationEvaluator(tokenizer, n, selected), ExtractScoresEvaluator(tokenizer, n, selected)]
    out = io.StringIO()
    w = writer_cls(evaluators)
    w.init_file(out)
    w.write_evaluation(0, batch, logits)
    return out.getvalue()


expected = {"tokens": tokens, "num_predictions": n, "logits": logits.tolist(), "batch": {k: v.tolist() for k, v in batch.items()},
            "multi_line": run(CSVMultiLineWriter), "single_line": run(CSVSingleLineWriter)}
with open(os.path.join(HERE, "predict_small.json"), "w") as f:
    json.dump(expected, f, indent=1)
print(expected["multi_line"][:400])
print(expected["single_line"][:400])
