"""Drop-in check: build every registered module key through the UNMODIFIED reference's own
container -- ``load_config`` -> ``create_container`` (``ImportFactory`` loads ``asme_b200.plugin`` from the ``imports:`` section,
``GenericModuleFactory`` / ``GenericModelFactory`` hand the constructors ``None`` for ``item_vocab_size`` / ``item_tokenizer`` /
``additional_attributes_tokenizer``, ``MetricsContainerFactory`` builds the reference's metric containers) -- on the reference's
``tests/example_dataset``, with the ``module`` sections of ``configs-new/{bert4rec,sasrec-cross,sasrec-neg}/ml-1m.yaml`` (paths and
column names pointed at the example data, negatives per user cut to what a 13-item vocabulary allows) and hand-written sections
for the keys ``configs-new`` has no usable file for (kebert4rec: its only file predates the model's ``prefusion_attributes``
argument; ubert4rec / user-sasrec-full: none).

Prints one JSON object {key: {...what was built...}}.  Runs in the build container only (needs /root/reference); no GPU: nothing
is executed beyond construction, the data loaders and the batch -> model-input plumbing.

    python tests/golden/_shims/build_through_container.py [key ...]
"""
import copy
import json
import os
import sys
import tempfile
from pathlib import Path

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ref_shims.install()
import yaml  # noqa: E402

PLUGIN_IMPORT = {"asme_b200": {"path": os.path.join(REPO, "recsys-22-user-attributes-recommender_b200"), "module": "asme_b200.plugin"}}

FEATURE = {"sequence_length": 200, "tokenizer": {"special_tokens": {"pad_token": "<PAD>", "mask_token": "<MASK>", "unk_token": "<UNK>"},
                                                 "vocabulary": None}}
SMALL_MODEL = {"max_seq_length": 20, "num_transformer_heads": 2, "num_transformer_layers": 1, "transformer_hidden_size": 16,
               "transformer_dropout": 0.1}
HAND_WRITTEN = {
    "kebert4rec": {"type": "kebert4rec", "metrics": {"full": {"metrics": {"recall": [1, 5], "ndcg": [5]}}},
                   "model": dict(SMALL_MODEL, prefusion_attributes={"attr_one": {"embedding_type": "content_embedding"}})},
    "ubert4rec": {"type": "ubert4rec", "metrics": {"full": {"metrics": {"recall": [1, 5], "mrr": [5]}}},
                  "model": dict(SMALL_MODEL, additional_attributes={"attr_one": {"embedding_type": "content_embedding"}},
                                user_attributes={"user_id": {"embedding_type": "user_embedding"}})},
    "user-sasrec-full": {"type": "user-sasrec-full", "metrics": {"full": {"metrics": {"recall": [1, 5], "mrr": [5]}}},
                         "model": dict(SMALL_MODEL, mode="full", user_attributes={"user_id": {"embedding_type": "user_embedding"}})},
}
REFERENCE_FILE = {"bert4rec": "bert4rec/ml-1m.yaml", "sasrec-cross": "sasrec-cross/ml-1m.yaml", "sasrec-neg": "sasrec-neg/ml-1m.yaml"}


def config_for(key, workdir):
    base = yaml.safe_load(open("/root/reference/configs-new/sasrec-cross/ml-1m.yaml"))
    if key in REFERENCE_FILE:
        ref = yaml.safe_load(open(os.path.join("/root/reference/configs-new", REFERENCE_FILE[key])))
        module = ref["module"]
        sources = ref["datamodule"]["data_sources"]
        sampled = module["metrics"].get("sampled")
        if sampled is not None:
            sampled["sample_probability_file"] = "example.popularity.item_id.txt"
            sampled["num_negative_samples"] = 2
    else:
        module = copy.deepcopy(HAND_WRITTEN[key])
        sources = copy.deepcopy(base["datamodule"]["data_sources"])
        if key in ("kebert4rec", "ubert4rec"):
            sources["train"]["processors"] = [{"type": "cloze", "mask_probability": 0.2, "only_last_item_mask_prob": 0.1}]
            for part in ("validation", "test"):
                sources[part]["processors"] = [{"type": "target_extractor"}, {"type": "last_item_mask"}]
    sources.update(file_prefix="example", num_workers=0, batch_size=4)
    features = {"item": dict(copy.deepcopy(FEATURE), column_name="item_id")}
    names = set(module["model"].get("prefusion_attributes") or {}) | set(module["model"].get("additional_attributes") or {}) | \
        set(module["model"].get("user_attributes") or {})
    for name in sorted(names):
        features[name] = dict(copy.deepcopy(FEATURE), column_name=name)
    cfg = {"imports": PLUGIN_IMPORT,
           "datamodule": {"dataset": "example", "data_sources": sources,
                          "preprocessing": {"input_file_path": "/root/reference/tests/example_dataset/example.csv",
                                            "output_directory": os.path.join(workdir, "data")}},
           "templates": {"unified_output": {"path": os.path.join(workdir, "out")}},
           "module": module, "features": features,
           "trainer": {"loggers": {"tensorboard": None}, "checkpoint": base["trainer"]["checkpoint"], "gpus": 0, "max_epochs": 1}}
    return cfg


def describe(key, container):
    module = container.module()
    model = module.model
    tok = module.item_tokenizer
    info = {"module_class": f"{type(module).__module__}.{type(module).__name__}",
            "model_class": f"{type(model).__module__}.{type(model).__name__}",
            "is_lightning_module": any(c.__name__ == "LightningModule" for c in type(module).__mro__),
            "item_vocab_size": model.item_vocab_size, "tokenizer_len": len(tok),
            "tokenizer_class": f"{type(tok).__module__}.{type(tok).__name__}",
            "pad": tok.pad_token_id, "mask": tok.mask_token_id,
            "metrics_class": f"{type(module.metrics).__module__}.{type(module.metrics).__name__}",
            "metric_names": module.metrics.get_metric_names(),
            "required_metadata_keys": model.required_metadata_keys(), "optional_metadata_keys": model.optional_metadata_keys(),
            "table_rows": {name: tuple(p.shape) for name, p in model.named_parameters() if "embedding" in name and p.dim() == 2},
            "optimizer": type(module.configure_optimizers() if not isinstance(module.configure_optimizers(), (list, tuple))
                              else module.configure_optimizers()[0][0] if isinstance(module.configure_optimizers()[0], list)
                              else module.configure_optimizers()[0]).__name__}
    batch = next(iter(container.train_dataloader()))
    info["train_batch_keys"] = sorted(batch.keys())
    missing = [k for k in model.required_metadata_keys() if k not in batch]
    info["missing_metadata"] = missing
    return info


def main():
    keys = sys.argv[1:] or ["bert4rec", "kebert4rec", "sasrec-cross", "sasrec-neg", "ubert4rec", "user-sasrec-full"]
    from asme.core.utils.run_utils import create_container, load_config
    out = {}
    for key in keys:
        workdir = tempfile.mkdtemp(prefix=f"dropin_{key}_")
        cfg_path = os.path.join(workdir, "cfg.yaml")
        yaml.safe_dump(config_for(key, workdir), open(cfg_path, "w"))
        try:
            out[key] = describe(key, create_container(load_config(Path(cfg_path))))
        except BaseException as e:          # the reference exits the process when a plug-in fails to import
            import traceback
            out[key] = {"error": f"{type(e).__name__}: {e}", "traceback": traceback.format_exc()[-2000:]}
    sys.stdout.write("\nRESULT_JSON " + json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
