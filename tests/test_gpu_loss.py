"""The training losses of supervision/direct.py:3-27 on the GPU (omnifusion_b200.supervision.direct, csrc/loss.cu)
against the committed outputs of the real reference (tests/golden/loss.npz) and the numpy oracle, forward and
gradient, including the reference's NaN results for degenerate inputs.  float32; tolerance 2e-6 relative (the
reductions run in another order)."""
import os

import numpy as np
import pytest
import torch

from omnifusion_b200.supervision import direct
from oracle import loss as ol
from test_oracle_golden import _loss_cases, _nan_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", ["small", "erp", "outliers", "equal", "empty_sample"])
def test_berhu_and_l1_match_reference_golden(name, golden_dir):
    z = np.load(os.path.join(golden_dir, "loss.npz"))
    cases, make_inputs = _loss_cases()
    shape, seed, kind = cases[name]
    pred, gt, mask, weights = make_inputs(shape, seed, kind)
    p = pred.to(DEV).requires_grad_(True)
    loss = direct.calculate_berhu_loss(p, gt.to(DEV), mask.to(DEV), weights.to(DEV))
    loss.backward()
    assert loss.shape == () and loss.dtype == torch.float32
    print(f"[parity] BerHu {name}: loss {loss.item():.7g} (reference {float(z[name + '_berhu']):.7g})")
    assert _nan_close(loss.item(), z[f"{name}_berhu"], 2e-6, 0)
    assert _nan_close(p.grad.cpu().numpy(), z[f"{name}_berhu_grad"], 2e-6, 1e-10)
    ref, _ = ol.berhu_loss(pred.numpy(), gt.numpy(), mask.numpy(), weights.numpy())
    assert _nan_close(loss.item(), ref, 2e-6, 0)
    p1 = pred.to(DEV).requires_grad_(True)
    l1 = direct.calculate_l1_loss(p1, gt.to(DEV), mask.to(DEV))
    l1.backward()
    print(f"[parity] masked L1 {name}: loss {l1.item():.7g} (reference {float(z[name + '_l1']):.7g})")
    assert _nan_close(l1.item(), z[f"{name}_l1"], 2e-6, 0)
    assert _nan_close(p1.grad.cpu().numpy(), z[f"{name}_l1_grad"], 2e-6, 1e-10)


def test_berhu_at_training_size_and_without_grad():
    """One training batch of train_erp_depth_iterative.py:271 (4 x 512 x 1024) against the oracle; under no_grad the
    same value comes back without a graph; a scaled upstream gradient scales d/dpred."""
    g = torch.Generator().manual_seed(9)
    gt = 0.1 + 7.9 * torch.rand(4, 1, 512, 1024, generator=g)
    pred = gt + 0.3 * torch.randn(gt.shape, generator=g)
    mask = torch.rand(gt.shape, generator=g) > 0.2
    w = torch.ones_like(gt)
    ref, c = ol.berhu_loss(pred.numpy(), gt.numpy(), mask.numpy(), w.numpy())
    p = pred.to(DEV).requires_grad_(True)
    loss = direct.calculate_berhu_loss(p, gt.to(DEV), mask.to(DEV), w.to(DEV))
    (3.0 * loss).backward()
    with torch.no_grad():
        again = direct.calculate_berhu_loss(p, gt.to(DEV), mask.to(DEV), w.to(DEV))
    assert not again.requires_grad and again.item() == loss.item()
    print(f"[parity] BerHu 4x512x1024: loss {loss.item():.7g} oracle {float(ref):.7g} c {c:.5g}")
    assert abs(loss.item() - float(ref)) <= 2e-6 * abs(float(ref))
    want = 3.0 * ol.berhu_grad(pred.numpy(), gt.numpy(), mask.numpy(), w.numpy())
    assert np.allclose(p.grad.cpu().numpy(), want, rtol=2e-6, atol=1e-12)
    with pytest.raises(Exception):
        direct.calculate_berhu_loss(pred, gt, mask, w)                      # CPU tensors: no CPU path
