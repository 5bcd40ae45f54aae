"""Drop-in boundary: the plug-in builds through the UNMODIFIED reference's own container.

``tests/golden/_shims/build_through_container.py`` (a subprocess: it installs import shims for the packages the reference needs and
this image lacks) runs ``load_config`` -> ``create_container`` -> ``container.module()`` for all six registered keys with an
``imports:`` section naming ``asme_b200.plugin``; here the result is checked: B200 classes, vocabulary sizes / tokenizers injected
from the reference's build context (utils/inject.py:174-198), the reference-built metric containers re-expressed.
Needs /root/reference (build container); the injection helpers themselves are also tested without it."""
import json
import os
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HELPER = os.path.join(ROOT, "tests", "golden", "_shims", "build_through_container.py")
KEYS = ["bert4rec", "kebert4rec", "sasrec-cross", "sasrec-neg", "ubert4rec", "user-sasrec-full"]
MODULES = {"bert4rec": "MaskedTrainingModule", "kebert4rec": "MaskedTrainingModule", "sasrec-cross": "NextItemPredictionTrainingModule",
           "sasrec-neg": "SequenceNextItemPredictionTrainingModule", "ubert4rec": "UBERTMaskedTrainingModule",
           "user-sasrec-full": "UserNextItemPredictionTrainingModule"}
MODELS = {"bert4rec": "BERT4RecModel", "kebert4rec": "KeBERT4RecModel", "sasrec-cross": "SASRecModel", "sasrec-neg": "SASRecModel",
          "ubert4rec": "UBERT4RecModel", "user-sasrec-full": "UserSASRecModel"}


@pytest.fixture(scope="module")
def built():
    if not os.path.isdir("/root/reference/src/asme"):
        pytest.skip("the reference tree is only present in the build container")
    r = subprocess.run([sys.executable, HELPER] + KEYS, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT_JSON ")]
    assert lines, f"helper printed no result (rc {r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    return json.loads(lines[-1][len("RESULT_JSON "):])


@pytest.mark.parametrize("key", KEYS)
def test_key_builds_through_the_reference_container(built, key):
    info = built[key]
    assert "error" not in info, info.get("traceback")
    assert info["module_class"] == f"asme_b200.modules.{MODULES[key]}"
    assert info["model_class"] == f"asme_b200.models.{MODELS[key]}"
    assert info["is_lightning_module"]                      # pl.Trainer type-checks the module it is given
    # InjectVocabularySize("item") / InjectTokenizer("item"): the example dataset has 10 items + PAD / MASK / UNK
    assert info["item_vocab_size"] == info["tokenizer_len"] == 13
    assert info["tokenizer_class"] == "asme.core.tokenization.tokenizer.Tokenizer"
    assert (info["pad"], info["mask"]) == (0, 1)
    assert info["metrics_class"] == "asme_b200.metrics.AggregateMetricsContainer"
    assert info["optimizer"] == "FusedAdam"
    assert info["missing_metadata"] == []                   # every attribute the model asks for is in the reference's batches
    item_tables = [shape for name, shape in info["table_rows"].items() if name.endswith("item_embedding.embedding.weight")]
    assert item_tables and all(shape[0] == 13 for shape in item_tables)


def test_reference_metric_sections_are_reexpressed(built):
    names = built["sasrec-cross"]["metric_names"]
    full = [f"{m}@{k}" for m in ("MRR", "recall", "NDCG") for k in (1, 5, 10)]
    assert names == full + [n + "sampled" for n in full]       # configs-new/sasrec-cross/ml-1m.yaml: full + sampled sections


def test_attribute_vocabularies_come_from_the_injected_tokenizers(built):
    ke = built["kebert4rec"]["table_rows"]
    assert ke["_sequence_embedding_layer.prefusion_attribute_embeddings.attr_one.weight"][0] == 7        # 4 values + 3 specials
    ub = built["ubert4rec"]
    assert ub["table_rows"]["_sequence_embedding_layer.user_attribute_embeddings.user_id.weight"][0] == 10
    assert ub["optional_metadata_keys"] == ["user_id"] and "attr_one" in ub["required_metadata_keys"]
    assert built["sasrec-neg"]["train_batch_keys"] == ["item", "length", "negative_samples", "pos", "positive_samples", "sample_ids"]


# ---- the injection helpers without the reference ---------------------------------------------------------------------------------
class _Tok:
    pad_token_id, mask_token_id = 0, 1

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


class _Ctx:
    def __init__(self, d):
        self.d = d

    def get(self, key):
        return self.d.get(key)

    def as_dict(self):
        return dict(self.d)


def _with_fake_asme(monkeypatch, ctx):
    fac = types.ModuleType("asme.core.init.factories")
    fac.GLOBAL_ASME_INJECTION_CONTEXT = None if ctx is None else types.SimpleNamespace(get_context=lambda: ctx)
    monkeypatch.setitem(sys.modules, "asme.core.init.factories", fac)


def test_injection_helpers(monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200"))
    from asme_b200 import inject
    monkeypatch.delitem(sys.modules, "asme.core.init.factories", raising=False)
    assert inject.injection_context() is None
    assert inject.resolve_vocab_size("item", 7) == 7
    with pytest.raises(KeyError):
        inject.resolve_vocab_size("item", None)
    assert inject.resolve_tokenizer("item") is None and inject.resolve_tokenizers() is None
    _with_fake_asme(monkeypatch, None)                        # asme imported, container not built yet
    assert inject.injection_context() is None
    ctx = _Ctx({"tokenizers.item": _Tok(13), "tokenizers.attr": _Tok(5), "datamodule": object()})
    _with_fake_asme(monkeypatch, ctx)
    assert inject.resolve_vocab_size("item", None) == 13
    assert inject.resolve_vocab_size("item", 99) == 99       # explicit values win (documented difference to @inject)
    assert inject.resolve_tokenizer("attr").n == 5
    assert sorted(inject.resolve_tokenizers()) == ["tokenizers.attr", "tokenizers.item"]
    with pytest.raises(KeyError):
        inject.resolve_tokenizer("missing", required=True)


def test_models_take_sizes_from_the_context(monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200"))
    from asme_b200.models import BERT4RecModel, KeBERT4RecModel
    from asme_b200.modules import MaskedTrainingModule
    ctx = _Ctx({"tokenizers.item": _Tok(21), "tokenizers.genre": _Tok(6)})
    _with_fake_asme(monkeypatch, ctx)
    model = BERT4RecModel(transformer_hidden_size=16, num_transformer_heads=2, num_transformer_layers=1, item_vocab_size=None,
                          max_seq_length=8, transformer_dropout=0.0)
    assert model.item_vocab_size == 21 and model.state_dict()["_projection_layer.output_bias"].shape == (21,)
    ke = KeBERT4RecModel(16, 2, 1, None, 8, 0.0, prefusion_attributes={"genre": {"embedding_type": "content_embedding"}},
                         additional_attributes_tokenizer=None)
    assert ke.state_dict()["_sequence_embedding_layer.prefusion_attribute_embeddings.genre.weight"].shape == (6, 16)
    module = MaskedTrainingModule(model, item_tokenizer=None, metrics=None)
    assert module.item_tokenizer is ctx.d["tokenizers.item"]
