"""ASME plug-in: re-registers ``bert4rec``, ``kebert4rec``, ``sasrec-cross``, ``sasrec-neg``, ``ubert4rec`` and
``user-sasrec-full`` in the reference's
module registry with the B200 model / module classes (modules/registry.py:19-22 ``register_module(...,
overwrite=True)``; activated through the reference's own ``imports:`` hook,
init/factories/include/import_factory.py:51-81).  Add to any existing ASME config:

    imports:
      asme_b200:
        path: /path/to/repo/recsys-22-user-attributes-recommender_b200
        module: asme_b200.plugin

Importing this module outside an ASME process (asme not importable) is a no-op apart from exposing
``REGISTRATIONS``, so that the boundary can be tested without the reference installed.
"""
from .models import BERT4RecModel, KeBERT4RecModel, SASRecModel, UBERT4RecModel, UserSASRecModel
from .modules import (MaskedTrainingModule, NextItemPredictionTrainingModule, SequenceNextItemPredictionTrainingModule,
                      UBERTMaskedTrainingModule, UserNextItemPredictionTrainingModule)

REGISTRATIONS = {
    "bert4rec": (MaskedTrainingModule, BERT4RecModel),
    "kebert4rec": (MaskedTrainingModule, KeBERT4RecModel),
    "sasrec-cross": (NextItemPredictionTrainingModule, SASRecModel),
    "sasrec-neg": (SequenceNextItemPredictionTrainingModule, SASRecModel),
    "ubert4rec": (UBERTMaskedTrainingModule, UBERT4RecModel),                      # modules/config.py:33
    "user-sasrec-full": (UserNextItemPredictionTrainingModule, UserSASRecModel),   # modules/config.py:49
}


def register(verbose: bool = False) -> bool:
    """Returns True when the registrations were written into a live ASME registry."""
    try:
        from asme.core.init.factories.modules.modules import GenericModuleFactory
        from asme.core.modules.registry import ModuleConfig, register_module
    except Exception as e:  # asme (or one of its dependencies) is not importable here
        if verbose:
            print(f"[asme_b200] ASME registry not available ({type(e).__name__}: {e}); nothing registered")
        return False
    for key, (module_cls, model_cls) in REGISTRATIONS.items():
        register_module(key, ModuleConfig(GenericModuleFactory, module_cls, {"model_cls": model_cls}), overwrite=True)
    return True


REGISTERED = register()
