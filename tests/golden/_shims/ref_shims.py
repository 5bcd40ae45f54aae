"""Import shims that let the UNMODIFIED reference (``/root/reference/src``) be imported in
the build container, where pytorch_lightning / torchmetrics / flatten_dict / dataclasses_json /
aim / optuna / _jsonnet / mlflow / redis are not installed.

Test infrastructure only: used by ``tests/golden/make_golden.py`` (which runs in the build
container, where ``/root/reference`` exists) to generate the committed golden fixtures. Nothing
here is imported by the product package, by ``-m gpu`` tests, by ``bench.py`` or by ``smoke()``.

The arithmetic of the reference path is stock torch; the shimmed packages contribute only
bookkeeping (``Metric.add_state``, ``LightningModule.log`` ...), see SURVEY.md Appendix A.
"""
import importlib.abc
import importlib.machinery
import random
import sys
import types

import numpy as np
import torch

REFERENCE_SRC = "/root/reference/src"


class _Placeholder:
    """Stands in for any class the reference subclasses or names at import time."""

    def __init__(self, *a, **k):
        pass

    def __init_subclass__(cls, **k):
        super().__init_subclass__()


class _FabricatedModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        value = type(name, (_Placeholder,), {})
        setattr(self, name, value)
        return value


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("pytorch_lightning", "aim", "optuna", "_jsonnet", "mlflow", "redis", "loguru")

    def find_spec(self, fullname, path, target=None):
        root = fullname.split(".")[0]
        if root in self.PREFIXES and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _FabricatedModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        pass


def _make_pytorch_lightning():
    pl = _FabricatedModule("pytorch_lightning")
    pl.__path__ = []

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, name, value, *a, **k):
            if not hasattr(self, "_logged"):
                object.__setattr__(self, "_logged", {})
            self._logged[name] = value

    class Callback:
        pass

    class Trainer:
        def __init__(self, *a, **k):
            pass

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    def seed_everything(seed, *a, **k):
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        return seed

    pl.LightningModule = LightningModule
    pl.Callback = Callback
    pl.Trainer = Trainer
    pl.LightningDataModule = LightningDataModule
    pl.seed_everything = seed_everything
    core = _FabricatedModule("pytorch_lightning.core")
    core.__path__ = []
    core.LightningDataModule = LightningDataModule
    core.LightningModule = LightningModule
    pl.core = core
    sys.modules["pytorch_lightning"] = pl
    sys.modules["pytorch_lightning.core"] = core


def _make_torchmetrics():
    tm = types.ModuleType("torchmetrics")
    tm.__path__ = []
    metric_mod = types.ModuleType("torchmetrics.metric")
    util_mod = types.ModuleType("torchmetrics.utilities")

    class Metric(torch.nn.Module):
        def __init__(self, compute_on_step=True, dist_sync_on_step=False, process_group=None, dist_sync_fn=None):
            super().__init__()
            self._defaults = {}

        def add_state(self, name, default, dist_reduce_fx=None):
            self._defaults[name] = default
            setattr(self, name, default.clone() if torch.is_tensor(default) else list(default))

        def reset(self):
            for name, default in self._defaults.items():
                setattr(self, name, default.clone() if torch.is_tensor(default) else list(default))

        def forward(self, *a, **k):
            # real torchmetrics returns the batch-local value while accumulating the global state
            saved = {n: (getattr(self, n).clone() if torch.is_tensor(getattr(self, n)) else list(getattr(self, n)))
                     for n in self._defaults}
            self.reset()
            self.update(*a, **k)
            batch_value = self.compute()
            for n, old in saved.items():
                cur = getattr(self, n)
                setattr(self, n, old + cur)
            return batch_value

    def reduce(x, reduction):
        if reduction == "elementwise_mean":
            return torch.mean(x)
        if reduction == "sum":
            return torch.sum(x)
        if reduction in (None, "none"):
            return x
        raise ValueError(reduction)

    metric_mod.Metric = Metric
    tm.Metric = Metric
    tm.metric = metric_mod
    util_mod.reduce = reduce
    tm.utilities = util_mod
    sys.modules["torchmetrics"] = tm
    sys.modules["torchmetrics.metric"] = metric_mod
    sys.modules["torchmetrics.utilities"] = util_mod


def _make_small():
    fd = types.ModuleType("flatten_dict")

    def flatten(d, reducer=None, **k):
        out = {}

        def rec(prefix, node):
            if isinstance(node, dict) and node:
                for key, val in node.items():
                    rec(prefix + (key,), val)
            else:
                out[prefix] = node

        rec((), d)
        return out

    fd.flatten = flatten
    sys.modules["flatten_dict"] = fd

    dj = types.ModuleType("dataclasses_json")
    dj.dataclass_json = lambda cls=None, **k: cls if cls is not None else (lambda c: c)
    sys.modules["dataclasses_json"] = dj


def install():
    """Install the shims and put the reference sources on ``sys.path`` (idempotent)."""
    if getattr(install, "_done", False):
        return
    _make_pytorch_lightning()
    _make_torchmetrics()
    _make_small()
    sys.meta_path.insert(0, _Finder())
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    install._done = True


def make_tokenizer(vocab_size, prefix="item"):
    """A reference ``Tokenizer`` with <PAD>=0, <MASK>=1, <UNK>=2 and ``vocab_size-3`` items."""
    from collections import OrderedDict

    from asme.core.tokenization.tokenizer import Tokenizer
    from asme.core.tokenization.vocabulary import Vocabulary

    tokens = OrderedDict()
    tokens["<PAD>"] = 0
    tokens["<MASK>"] = 1
    tokens["<UNK>"] = 2
    for i in range(3, vocab_size):
        tokens[f"{prefix}_{i}"] = i
    return Tokenizer(Vocabulary(tokens), pad_token="<PAD>", mask_token="<MASK>", unk_token="<UNK>")


_CTX = {}


def set_injection_context(tokenizers):
    """Bind tokenizers (``{"item": tok, "<attr>": tok}``) into the reference's global injection
    context (the global is copied at import time, so the same Context is mutated)."""
    import asme.core.init.factories as factories
    from asme.core.init.config import Config
    from asme.core.init.context import Context

    if "ctx" not in _CTX:
        ctx = Context()
        _CTX["ctx"] = ctx
        factories.GLOBAL_ASME_INJECTION_CONTEXT = factories.BuildContext(Config({}), ctx)
        inj = sys.modules.get("asme.core.utils.inject")
        if inj is not None:
            inj.GLOBAL_ASME_INJECTION_CONTEXT = factories.GLOBAL_ASME_INJECTION_CONTEXT
    ctx = _CTX["ctx"]
    for name, tok in tokenizers.items():
        ctx.set(f"tokenizers.{name}", tok, overwrite=True)
    return ctx
