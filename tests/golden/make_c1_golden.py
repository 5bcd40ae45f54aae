"""C1 (BASELINE.json configs[0]): the UNMODIFIED reference runner on its own ``tests/example_dataset`` with the module / model
sections of ``configs-new/sasrec-cross/ml-1m.yaml`` (SURVEY.md 8c recipe), driven for a few epochs by a hook loop (Lightning is
absent here): ``load_config`` -> ``create_container`` (preprocessing, tokenizers, ``NextItemPredictionTrainingModule(SASRecModel)``,
metrics, the three dataloaders) -> training_step / backward / Adam, validation_step -> validation_step_end ->
validation_epoch_end.

    python tests/golden/make_c1_golden.py

Writes ``c1_sasrec_cross_example.npz``: the module's initial state dict (``w::<name>``), every training batch it saw
(``train::<epoch>::<i>::item|target``) with its loss, every validation / test batch with its ``val_loss`` and step metrics, the
per-epoch ``recall@k`` / ``NDCG@k`` / ``MRR@k`` values, and ``c1_config.yaml`` = the exact config the run used (the reference's
file with the dataset paths pointed at the example data, dropout 0, no sampled metrics, batch size 4).

Runs in the build container only (needs /root/reference); the GPU test ``tests/test_gpu_c1.py`` replays the recorded batches.
"""
import json
import os
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shims"))
import ref_shims  # noqa: E402

ref_shims.install()

EPOCHS = 3
BATCH = 4


def build_config(workdir):
    cfg = yaml.safe_load(open("/root/reference/configs-new/sasrec-cross/ml-1m.yaml"))
    dm = cfg["datamodule"]
    dm["dataset"] = "example"
    dm["data_sources"]["file_prefix"] = "example"
    dm["data_sources"]["num_workers"] = 0
    dm["data_sources"]["batch_size"] = BATCH
    dm["preprocessing"] = {"input_file_path": "/root/reference/tests/example_dataset/example.csv",
                           "output_directory": os.path.join(workdir, "data")}
    cfg["features"]["item"]["column_name"] = "item_id"
    cfg["module"]["model"]["transformer_dropout"] = 0.0
    cfg["module"]["metrics"].pop("sampled")          # needs the ml-1m popularity file
    cfg["trainer"]["gpus"] = 0
    cfg["trainer"]["loggers"] = {"tensorboard": None}
    cfg["trainer"]["max_epochs"] = EPOCHS
    cfg["templates"]["unified_output"]["path"] = os.path.join(workdir, "out")
    return cfg


def main():
    workdir = tempfile.mkdtemp(prefix="c1_")
    cfg = build_config(workdir)
    cfg_path = os.path.join(workdir, "cfg.yaml")
    yaml.safe_dump(cfg, open(cfg_path, "w"))
    from asme.core.utils.run_utils import create_container, load_config
    torch.manual_seed(0)
    container = create_container(load_config(Path(cfg_path)))
    module = container.module()
    out = {}
    for name, v in module.state_dict().items():
        out[f"w::{name}"] = v.detach().clone().numpy()
    optimizer = module.configure_optimizers()
    summary = {"epochs": []}
    torch.manual_seed(1)          # shuffling order of the training loader
    for epoch in range(EPOCHS):
        module.train()
        losses = []
        for i, batch in enumerate(container.train_dataloader()):
            out[f"train::{epoch}::{i}::item"] = batch["item"].numpy()
            out[f"train::{epoch}::{i}::target"] = batch["item.target"].numpy()
            optimizer.zero_grad()
            loss = module.training_step(batch, i)["loss"]
            loss.backward()
            optimizer.step()
            losses.append(float(loss))
        module.eval()
        val_losses, step_values = [], []
        with torch.no_grad():
            for i, batch in enumerate(container.validation_dataloader()):
                out[f"val::{epoch}::{i}::item"] = batch["item"].numpy()
                out[f"val::{epoch}::{i}::target"] = batch["item.target"].numpy()
                res = module.validation_step(batch, i)
                val_losses.append(float(module._logged["val_loss"]))
                step_values.append({k: float(v) for k, v in module.validation_step_end(res).items()})
            module.validation_epoch_end([])
        epoch_metrics = {k: float(v) for k, v in module._logged.items() if "@" in k}
        summary["epochs"].append({"train_loss": losses, "val_loss": val_losses, "val_step": step_values, "val_epoch": epoch_metrics})
        print(epoch, losses, epoch_metrics)
    # test pass with the final weights
    with torch.no_grad():
        tl, ts = [], []
        for i, batch in enumerate(container.test_dataloader()):
            out[f"test::{i}::item"] = batch["item"].numpy()
            out[f"test::{i}::target"] = batch["item.target"].numpy()
            res = module.test_step(batch, i)
            ts.append({k: float(v) for k, v in module.test_step_end(res).items()})
        module.test_epoch_end([])
        summary["test_step"] = ts
        summary["test_epoch"] = {k: float(v) for k, v in module._logged.items() if "@" in k}
    for name, v in module.state_dict().items():
        out[f"final::{name}"] = v.detach().clone().numpy()
    tok = module.item_tokenizer
    summary["tokenizer"] = {"len": len(tok), "pad": tok.pad_token_id, "mask": tok.mask_token_id, "unk": tok.unk_token_id}
    summary["module"] = {"learning_rate": module.learning_rate, "beta_1": module.beta_1, "beta_2": module.beta_2,
                         "weight_decay": module.weight_decay}
    summary["model"] = cfg["module"]["model"]
    summary["metrics"] = cfg["module"]["metrics"]
    out["summary_json"] = np.frombuffer(json.dumps(summary).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "c1_sasrec_cross_example.npz"), **out)
    # the config with the temporary paths replaced by placeholders
    cfg["datamodule"]["preprocessing"]["output_directory"] = "<tmp>/data"
    cfg["templates"]["unified_output"]["path"] = "<tmp>/out"
    yaml.safe_dump(cfg, open(os.path.join(HERE, "c1_config.yaml"), "w"))
    shutil.rmtree(workdir, ignore_errors=True)
    print("wrote c1_sasrec_cross_example.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
