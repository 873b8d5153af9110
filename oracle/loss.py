"""CPU restatement of the reference's training losses (supervision/direct.py:3-27).  TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py).  Pinned against outputs of the reference itself: tests/golden/make_golden_loss.py ->
tests/golden/loss.npz, checked by tests/test_oracle_golden.py."""
import numpy as np


def berhu_loss(pred, gt, mask, weights):
    """supervision/direct.py:3-18 in numpy float32.  c = max|gt - pred| / 5 over the whole unmasked batch is a Python
    float (double) that meets the float32 arrays as a float32 scalar, like torch's scalar promotion."""
    pred, gt, weights = (np.asarray(a, dtype=np.float32) for a in (pred, gt, weights))
    bs = pred.shape[0]
    diff = gt - pred
    abs_diff = np.abs(diff)
    c = float(np.max(abs_diff)) / 5                              # :7
    with np.errstate(divide="ignore", invalid="ignore"):
        leq = (abs_diff <= np.float32(c)).astype(np.float32)     # :8
        l2 = (diff * diff + np.float32(c ** 2)) / np.float32(2 * c)   # :9
        loss = leq * abs_diff + (1 - leq) * l2                   # :10
        loss = loss.reshape(bs, -1)
        m = np.asarray(mask).reshape(bs, -1).astype(np.float32)
        w = weights.reshape(bs, -1)
        count = m.sum(axis=1, keepdims=True, dtype=np.float32)   # :15
        per = (loss * m * w).sum(axis=1, keepdims=True, dtype=np.float32) / count
        return np.float32(per.mean(dtype=np.float32)), c


def berhu_grad(pred, gt, mask, weights):
    """d loss / d pred of the expression above with c held constant (the reference's .item())."""
    pred, gt, weights = (np.asarray(a, dtype=np.float32) for a in (pred, gt, weights))
    bs = pred.shape[0]
    diff = gt - pred
    abs_diff = np.abs(diff)
    c = float(np.max(abs_diff)) / 5
    with np.errstate(divide="ignore", invalid="ignore"):
        leq = (abs_diff <= np.float32(c)).astype(np.float32)
        dl = leq * (-np.sign(diff)) + (1 - leq) * ((np.float32(-2) * diff) / np.float32(2 * c))
        m = np.asarray(mask).astype(np.float32)
        count = m.reshape(bs, -1).sum(axis=1, dtype=np.float32).reshape((bs,) + (1,) * (pred.ndim - 1))
        return (dl * m * weights / count / np.float32(bs)).astype(np.float32)


def l1_loss(pred, gt, mask):
    """supervision/direct.py:20-27."""
    pred, gt = (np.asarray(a, dtype=np.float32) for a in (pred, gt))
    bs = pred.shape[0]
    loss = np.abs(gt - pred).reshape(bs, -1)
    m = np.asarray(mask).reshape(bs, -1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        per = (loss * m).sum(axis=1, dtype=np.float32) / m.sum(axis=1, dtype=np.float32)
        return np.float32(per.mean(dtype=np.float32))
