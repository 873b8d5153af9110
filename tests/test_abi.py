"""The C-ABI shared library loads and exports every symbol include/ofb.h declares
(no compute calls: this runs without a GPU), and the host-side pieces agree with it."""
import ctypes
import os
import re

import pytest
import torch

from omnifusion_b200 import _lib
from omnifusion_b200.checkpoint import key_spec, strip_module_prefix, synthetic_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ofb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ofb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libofb.so missing: run `make -C omnifusion_b200/csrc`"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/ofb.h but not exported"
    assert sorted(_lib._SIGNATURES) == syms, "python binding and header disagree on the symbol set"
    assert _lib.lib().ofb_version() == 103


def test_errors_are_reported_not_swallowed():
    L = _lib.lib()
    # null pointers are rejected before any CUDA call
    assert L.ofb_equi2pers_f32(None, 1, 3, 8, 16, None, 18, 4, 4, None, 0, None) < 0
    assert b"null" in L.ofb_last_error()
    with pytest.raises(_lib.OfbError):
        _lib.check(L.ofb_layernorm_f32(None, None, None, 1, 512, 1e-5, None, 0, 0, None))
    with pytest.raises(_lib.OfbError):
        _lib.require_cuda(torch.zeros(1, 3, 8, 16), "x")          # no CPU fallback
    # the newer entry points validate before they touch the device too
    assert L.ofb_depth_loss_f32(None, None, None, None, 2, 16, 1, None, None, None, None) < 0
    assert L.ofb_token_stack_f32(None, None, None, 0, None, 1, 18, 6, 0, None) < 0
    assert L.ofb_splitk_finish_conv_f16(None, 2, 16, 512, None, None, 1.0, None, 1, None, None) < 0
    assert L.ofb_loss_work_bytes(4) >= 16 and L.ofb_token_stack_scratch_floats(8, 18) > 8 * 18 * 512 * 17
    from omnifusion_b200.supervision import direct
    x = torch.zeros(2, 1, 4, 8)
    with pytest.raises(_lib.OfbError):
        direct.calculate_berhu_loss(x, x, x > 0, x)                # CPU tensors: the losses have no CPU path either


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "omnifusion_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_key_spec_and_synthetic_checkpoint():
    for kind, n in (("iterative", 375), ("single", 363)):
        spec = key_spec(kind, 18)
        assert len(spec) == n
        sd = synthetic_state_dict(kind, 18, 0)
        assert list(sd) == list(spec)
        assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in spec)
    a, b = synthetic_state_dict("iterative", 18, 0), synthetic_state_dict("iterative", 18, 0)
    assert all(torch.equal(a[k], b[k]) for k in a)
    c = synthetic_state_dict("iterative", 46, 0)
    assert c["transformer.pos_emb"].shape == (1, 46, 512) and torch.equal(c["conv1.weight"], a["conv1.weight"])
    pref = {"module." + k: v for k, v in a.items()}
    assert list(strip_module_prefix(pref)) == list(a)
    nparams = sum(v.numel() for k, v in a.items() if v.dtype == torch.float32 and "running" not in k)
    assert nparams == 42479042 - 0 or nparams > 42_000_000


def test_module_state_dict_matches_reference_layout():
    from omnifusion_b200.model.spherical_model import spherical_fusion as single
    from omnifusion_b200.model.spherical_model_iterative import spherical_fusion as iterative
    net = iterative(6, 46, (128, 128), (80, 80))
    assert list(net.state_dict()) == list(key_spec("iterative", 46))
    assert list(single().state_dict()) == list(key_spec("single", 18))
    with pytest.raises(ValueError):
        iterative(4, 26)
    with pytest.raises(_lib.OfbError):
        net(torch.zeros(1, 3, 64, 128), iter=1)                     # CPU tensors are refused
