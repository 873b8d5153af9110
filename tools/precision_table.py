#!/usr/bin/env python
"""Per-layer-class operand precision study on the CPU oracle (verdict r1 item 4): which layer classes tolerate a
single TF32 pass (operands rounded to 10 mantissa bits, fp32 accumulation) instead of the split-half three-product
scheme, measured as the max relative error of the final 2-iteration depth against the fp32 forward, on several
synthetic checkpoints.  Emulation: conv / linear inputs and weights of the selected class are rounded to TF32
(round-to-nearest-even, as cvt.rna would) before the fp32 op; everything else stays fp32.

    python tools/precision_table.py > profiles/r02_precision_table.md
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from omnifusion_b200.checkpoint import synthetic_state_dict
from oracle import model as om


def tf32(x):
    i = x.contiguous().view(torch.int32)
    r = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF          # round to nearest even on the 13 dropped bits
    return r.view(torch.float32)


def bf16x(x):
    return x.to(torch.bfloat16).to(torch.float32)


CLASSES = {
    "stem (7x7, 3->64)": lambda n: n == "conv1",
    "layer1 (64ch @32)": lambda n: n.startswith("layer1."),
    "layer2 (128ch @16)": lambda n: n.startswith("layer2."),
    "layer3 (256ch @8)": lambda n: n.startswith("layer3."),
    "layer4 (512ch @4)": lambda n: n.startswith("layer4."),
    "token path (down1 + 6 blocks)": lambda n: n.startswith("down") or n.startswith("transformer."),
    "decoder 0_0 .. 2_1 (<= 32x32)": lambda n: any(n.startswith(f"de_conv{i}_") for i in (0, 1, 2)),
    "decoder 3_0 / 3_1 (64x64)": lambda n: n.startswith("de_conv3_"),
    "de_conv4_0 + heads (128x128)": lambda n: n.startswith("de_conv4_") or n in ("pred", "weight_pred"),
    "ALL layers": lambda n: True,
}

_sel = [None]
_round = [tf32]
_orig_conv, _orig_linear = om._conv, F.linear


def conv_hook(sd, name, x, stride=1, pad=0):
    if _sel[0] and _sel[0](name):
        w = _round[0](om._w2d(sd, name + ".weight"))
        return F.conv2d(_round[0](x), w, sd.get(name + ".bias"), stride, pad)
    return _orig_conv(sd, name, x, stride, pad)


class LinearHook:
    """F.linear inside oracle.model.transformer / _attention: the weight tensor identifies the layer."""

    def __init__(self, sd):
        self.names = {id(v): k for k, v in sd.items()}

    def __call__(self, x, w, b=None):
        n = self.names.get(id(w), "")
        if _sel[0] and n and _sel[0](n):
            return _orig_linear(_round[0](x), _round[0](w), b)
        return _orig_linear(x, w, b)


def run(sd, rgb):
    om._conv = conv_hook
    om.F.linear = LinearHook(om.strip_module_prefix(sd))
    try:
        return om.forward_iterative(sd, rgb, 2, True)[-1]
    finally:
        om._conv = _orig_conv
        om.F.linear = _orig_linear


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    seeds = (0, 1, 2)
    rgb = torch.rand(1, 3, 64, 128, generator=torch.Generator().manual_seed(123))
    print("# Operand-precision sensitivity per layer class (CPU oracle emulation)\n")
    print("Max relative error of the final 2-iteration depth (B=1, 64x128 ERP, nrows=4, confidence on) against the fp32")
    print("forward when ONE layer class runs with TF32-rounded operands (one `kind::tf32` pass = 2 cost units) instead of")
    print("the split-half scheme (three `kind::f16` passes = 3 cost units, error ~1e-6 of fp32), on three synthetic")
    print("checkpoints (`omnifusion_b200.checkpoint.synthetic_state_dict(seed)`).  Bar: 1e-3 on the whole network;")
    print("the verdict's rule for dropping a class to one pass is <= 2e-4 alone.\n")
    print("| layer class | " + " | ".join(f"seed {s}" for s in seeds) + " | worst | <= 2e-4 |")
    print("|---|" + "---|" * (len(seeds) + 2))
    refs = {}
    rows = {}
    for s in seeds:
        sd = synthetic_state_dict("iterative", 18, s)
        _sel[0] = None
        refs[s] = (sd, om.forward_iterative(sd, rgb, 2, True)[-1])
    for rounding, label in ((tf32, "TF32"), (bf16x, "BF16")):
        _round[0] = rounding
        for cname, pred in CLASSES.items():
            if rounding is bf16x and cname != "ALL layers":
                continue
            errs = []
            for s in seeds:
                sd, ref = refs[s]
                _sel[0] = pred
                out = run(sd, rgb)
                _sel[0] = None
                errs.append(((out - ref).abs() / ref.abs().clamp_min(1e-6)).max().item())
            w = max(errs)
            print(f"| {cname} ({label}) | " + " | ".join(f"{e:.2e}" for e in errs) + f" | {w:.2e} | {'yes' if w <= 2e-4 else 'no'} |", flush=True)


if __name__ == "__main__":
    main()
