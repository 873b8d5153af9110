"""Import shim for the *real* reference (authoring container only).

Used by make_golden.py to generate the committed fixtures.  Never imported by
tests, smoke() or bench.py: /root/reference does not exist on the GPU box.

The reference imports timm / mmseg / mmcv at model/blocks.py:6-11 without using
them (DropPath is only built when drop_path > 0) and downloads ResNet-34 weights
at spherical_model_iterative.py:260; both are neutralised here.
"""
import os
import sys
import tempfile
import types

import torch.nn as nn
import torchvision

REF_ROOT = os.environ.get("OFB_REFERENCE_ROOT", "/root/reference")


def install():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("timm"); mod("timm.models")
    mod("timm.models.layers", DropPath=nn.Identity, to_2tuple=lambda x: (x, x),
        trunc_normal_=nn.init.trunc_normal_)
    mod("timm.models.registry", register_model=lambda f: f)
    mod("timm.models.vision_transformer", _cfg=lambda **kw: kw)
    mod("mmseg"); mod("mmseg.utils", get_root_logger=lambda *a, **k: None)
    mod("mmcv"); mod("mmcv.runner", load_checkpoint=lambda *a, **k: None)

    real = torchvision.models.resnet34
    torchvision.models.resnet34 = lambda pretrained=False, **kw: real(weights=None)

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def fresh_cwd():
    """The reference caches ./grid/<layer_name>.pth keyed by name only
    (pers2equi_v3.py:27); every config needs its own scratch cwd."""
    d = tempfile.mkdtemp(prefix="ofb_ref_")
    os.chdir(d)
    return d
