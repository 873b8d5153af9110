"""Depth metrics on the device, mirroring the reference's metrics.py:7-26 and the evaluation
step of test.py:151-177 (median scaling, N-weighted running means)."""
import torch

from . import _lib


def abs_rel_error(pred, gt, mask, scale=1.0):
    """metrics.py:7-9 via the ofb_absrel_partial reduction kernel. Returns (sum, count) float64 (2,)."""
    pred = _lib.require_cuda(pred, "pred")
    gt = _lib.require_cuda(gt, "gt")
    m = (mask > 0).to(torch.uint8).contiguous()
    out = torch.zeros(2, dtype=torch.float64, device=pred.device)
    _lib.use_device(pred.device)
    _lib.check(_lib.lib().ofb_absrel_partial(_lib.ptr(pred), _lib.ptr(gt), _lib.ptr(m), pred.numel(),
                                             float(scale), _lib.ptr(out), _lib.stream_of(pred.device)))
    return out


def median_scale(pred, gt, mask):
    """test.py:161-162: median(gt[mask]) / median(pred[mask]) over the whole batch tensor."""
    m = mask > 0
    return (gt[m].median() / pred[m].median()).item()


METRIC_NAMES = ("abs_rel", "sq_rel", "rms_sq_lin", "rms_sq_log", "d1", "d2", "d3")


def depth_metrics_partial(pred, gt, mask, scale=1.0):
    """All seven metrics of metrics.py:7-26 as float64 partial sums (9,) from one kernel pass:
    [abs_rel, sq_rel, rms_sq_lin, rms_sq_log sums, n_log, d1, d2, d3 counts, n]."""
    pred = _lib.require_cuda(pred, "pred")
    gt = _lib.require_cuda(gt, "gt")
    m = (mask > 0).to(torch.uint8).contiguous()
    out = torch.zeros(9, dtype=torch.float64, device=pred.device)
    _lib.use_device(pred.device)
    _lib.check(_lib.lib().ofb_depth_metrics_partial(_lib.ptr(pred), _lib.ptr(gt), _lib.ptr(m), pred.numel(),
                                                    float(scale), _lib.ptr(out), _lib.stream_of(pred.device)))
    return out


def finalize_metrics(partial):
    """(9,) partial sums (possibly all-reduced over ranks / batches) -> dict like test.py's meters."""
    p = partial.tolist()
    n = max(p[8], 1.0)
    return {"abs_rel": p[0] / n, "sq_rel": p[1] / n, "rms_sq_lin": p[2] / n, "rms_sq_log": p[3] / max(p[4], 1.0),
            "d1": p[5] / n, "d2": p[6] / n, "d3": p[7] / n, "n": int(p[8])}


def compute_eval_metrics(pred, gt, mask, use_median_scale=True):
    """test.py:151-170 for one batch tensor: median scaling, then the seven metrics."""
    s = median_scale(pred, gt, mask) if use_median_scale else 1.0
    return finalize_metrics(depth_metrics_partial(pred, gt, mask, s))


class AbsRelMeter:
    """N-weighted running mean like test.py's AverageMeter.update(val, N) (test.py:121-148), kept as
    (sum of per-batch mean * n, sum of n) so shards can be combined with one all-reduce(SUM)."""

    def __init__(self, device):
        self.acc = torch.zeros(2, dtype=torch.float64, device=device)

    def update(self, pred, gt, mask, use_median_scale=True):
        s = median_scale(pred, gt, mask) if use_median_scale else 1.0
        part = abs_rel_error(pred, gt, mask, s)          # (sum, n): mean*n == sum
        self.acc += part
        return (part[0] / part[1]).item()

    def all_reduce(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM)
        return self

    @property
    def avg(self):
        return (self.acc[0] / self.acc[1]).item()
