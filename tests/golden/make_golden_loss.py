"""Golden vectors for the training losses, produced by the REAL reference (supervision/direct.py imported from
/root/reference, which only needs torch).  Run in the authoring container:

    python tests/golden/make_golden_loss.py

Writes tests/golden/loss.npz: per case the inputs' generator seeds / shapes and the reference's loss and
d loss / d pred.  Cases: ragged masks, non-uniform weights, a large-outlier case (most pixels in the L2 branch), the
degenerate pred == gt case (c = 0 -> NaN in the reference) and a sample without valid pixels (0 / 0)."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_direct", "/root/reference/supervision/direct.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = {
    # name: (shape, seed, kind)
    "small": ((2, 1, 16, 32), 1, "plain"),
    "erp": ((3, 1, 64, 128), 2, "plain"),
    "outliers": ((2, 1, 32, 64), 3, "outliers"),
    "equal": ((2, 1, 8, 16), 4, "equal"),
    "empty_sample": ((2, 1, 8, 16), 5, "empty"),
}


def make_inputs(shape, seed, kind):
    g = torch.Generator().manual_seed(seed)
    gt = 0.1 + 7.9 * torch.rand(shape, generator=g)
    pred = gt + 0.5 * torch.randn(shape, generator=g)
    if kind == "outliers":
        pred = gt + 0.05 * torch.randn(shape, generator=g)
        pred.view(-1)[::97] += 6.0
    if kind == "equal":
        pred = gt.clone()
    mask = torch.rand(shape, generator=g) > 0.3
    if kind == "empty":
        mask[1] = False
    weights = 0.5 + torch.rand(shape, generator=g)
    return pred, gt, mask, weights


def main():
    out = {}
    for name, (shape, seed, kind) in CASES.items():
        pred, gt, mask, weights = make_inputs(shape, seed, kind)
        p = pred.clone().requires_grad_(True)
        loss = ref.calculate_berhu_loss(p, gt, mask, weights)
        loss.backward()
        p1 = pred.clone().requires_grad_(True)
        l1 = ref.calculate_l1_loss(p1, gt, mask)
        l1.backward()
        out[f"{name}_berhu"] = np.float32(loss.item())
        out[f"{name}_berhu_grad"] = p.grad.numpy()
        out[f"{name}_l1"] = np.float32(l1.item())
        out[f"{name}_l1_grad"] = p1.grad.numpy()
        print(name, loss.item(), l1.item())
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
