"""Constructor-argument injection, as the reference's ``@inject`` decorator performs it
(utils/inject.py:78-112, :174-205): the ASME factories hand model / module constructors ``None`` for every
``*_vocab_size`` and ``*_tokenizer`` parameter (init/factories/modules/modules.py:116-127, util.py:49-58) and the decorator
fills them from the build context -- ``len(context["tokenizers.<feature>"])``, the tokenizer itself, or all tokenizers.

The B200 classes resolve the same three kinds of values themselves, at CONSTRUCTION time, from the live global
``asme.core.init.factories.GLOBAL_ASME_INJECTION_CONTEXT`` (set by ``create_container``, utils/run_utils.py:101).  Nothing of
asme is imported here: the module is looked up in ``sys.modules`` (SURVEY.md Q8 -- importing ``asme.core.utils.inject`` before
the container exists would freeze a ``None`` context into it), so outside an ASME process everything below is a no-op and
explicit values are required.

One deliberate difference: the reference DISCARDS an explicitly passed value of an injected parameter (inject.py:107); here an
explicit non-None value wins, so the classes stay usable without a container (tests, bench, notebooks).
"""
import sys
from typing import Any, Dict, Optional

TOKENIZERS_PREFIX = "tokenizers"          # init/factories/features/tokenizer_factory.py:63-64


def injection_context():
    """the reference's live ``Context`` (string key -> object) or None"""
    factories = sys.modules.get("asme.core.init.factories")
    build_context = getattr(factories, "GLOBAL_ASME_INJECTION_CONTEXT", None) if factories is not None else None
    if build_context is None:
        return None
    return build_context.get_context()


def resolve_tokenizer(feature: str, given: Any = None, required: bool = False):
    """``InjectTokenizer(feature)`` (inject.py:176-186)"""
    if given is not None:
        return given
    ctx = injection_context()
    tok = ctx.get(f"{TOKENIZERS_PREFIX}.{feature}") if ctx is not None else None
    if tok is None and required:
        raise KeyError(f'No tokenizer with id "{feature}" configured and no default value set.')
    return tok


def resolve_vocab_size(feature: str, given: Optional[int] = None) -> int:
    """``InjectVocabularySize(feature)`` (inject.py:190-198)"""
    if given is not None:
        return int(given)
    tok = resolve_tokenizer(feature)
    if tok is None:
        raise KeyError(f"No tokenizer with id {feature} configured. Can not inject vocabulary size into parameter "
                       f"{feature}_vocab_size (pass it explicitly outside an ASME container).")
    return len(tok)


def resolve_tokenizers(given: Optional[Dict[str, Any]] = None) -> Optional[Dict[str, Any]]:
    """``InjectTokenizers()`` (inject.py:187-189 -> init/factories/modules/util.py:40-46): every tokenizer of the context under its
    context key (``"tokenizers.<feature>"``)"""
    if given is not None:
        return given
    ctx = injection_context()
    if ctx is None:
        return None
    return {key: obj for key, obj in ctx.as_dict().items()
            if key.startswith(TOKENIZERS_PREFIX + ".") and hasattr(obj, "pad_token_id") and hasattr(obj, "__len__")}
