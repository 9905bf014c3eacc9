"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/rsrx.h declares; struct sizes agree with the Python mirrors."""
import ctypes as C
import os
import re

import pytest

from rsr_mjx_b200 import _lib
from rsr_mjx_b200.model import EnvCfg, ModelBlob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "rsrx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rsrx_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_sizes_match(lib):
    assert lib.rsrx_model_blob_size() == C.sizeof(ModelBlob)
    assert lib.rsrx_env_cfg_size() == C.sizeof(EnvCfg)
    assert lib.rsrx_debug_stride() > 0
    assert b"sm_100a" in lib.rsrx_version()


def test_errors_are_return_codes_not_exceptions(lib):
    h = C.c_void_p()
    blob, cfg = ModelBlob(), EnvCfg()
    assert lib.rsrx_model_create(C.byref(blob), 12, C.byref(cfg), C.byref(h)) != 0
    assert b"blob size mismatch" in lib.rsrx_last_error()
    assert lib.rsrx_model_create(C.byref(blob), C.sizeof(blob), C.byref(cfg), C.byref(h)) != 0
    assert b"magic" in lib.rsrx_last_error()
    assert lib.rsrx_env_step(None, 4, _lib.StateC(), None, None, None) != 0
    assert b"null" in lib.rsrx_last_error()


def test_product_path_refuses_cpu():
    import torch
    from rsr_mjx_b200.envs import AirbotPlayBase
    from rsr_mjx_b200 import rsr_loss
    with pytest.raises(RuntimeError):
        AirbotPlayBase("sf", num_envs=2, device="cpu")
    with pytest.raises(RuntimeError):
        rsr_loss.evaluate_kde(torch.zeros(4, 3), torch.zeros(2, 3))
