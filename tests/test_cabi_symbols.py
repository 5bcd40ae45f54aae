"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports exactly the
symbols include/asme_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os

from asme_b200 import _lib


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _lib.header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/asme_b200.h but not exported"
    assert set(declared) == set(_lib._PROTOTYPES), "ctypes prototypes out of sync with the header"


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.asme_b200_abi_version() == 3
    # an invalid call fails loudly with a message, without touching the GPU
    rc = lib.asme_b200_layernorm_fwd(None, None, None, 4, 64, None, None, None, None)
    assert rc != 0
    assert b"null" in lib.asme_b200_last_error()


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from asme_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.layernorm_fwd(torch.randn(4, 64), torch.ones(64), torch.zeros(64))
