"""Depth metrics on the device, mirroring the reference's metrics.py:7-26 and the evaluation
step of test.py:151-177 (median scaling, N-weighted running means)."""
import torch

from . import _lib


def _same_shape(pred, gt, mask):
    # the kernels index all three tensors with one flat offset: broadcastable-but-different shapes would read
    # out of bounds
    if not (tuple(pred.shape) == tuple(gt.shape) == tuple(mask.shape)):
        raise ValueError(f"pred, gt and mask must have identical shapes, got {tuple(pred.shape)}, "
                         f"{tuple(gt.shape)}, {tuple(mask.shape)}")


def abs_rel_error(pred, gt, mask, scale=1.0):
    """metrics.py:7-9 via the ofb_absrel_partial reduction kernel. Returns (sum, count) float64 (2,)."""
    _same_shape(pred, gt, mask)
    pred = _lib.require_cuda(pred, "pred")
    gt = _lib.require_cuda(gt, "gt")
    m = (mask > 0).to(torch.uint8).contiguous()
    out = torch.zeros(2, dtype=torch.float64, device=pred.device)
    _lib.use_device(pred.device)
    _lib.check(_lib.lib().ofb_absrel_partial(_lib.ptr(pred), _lib.ptr(gt), _lib.ptr(m), pred.numel(),
                                             float(scale), _lib.ptr(out), _lib.stream_of(pred.device)))
    return out


def median_scale_device(pred, gt, mask):
    """test.py:161-162 on the device: [median(gt[mask]) / median(pred[mask]), median(gt[mask]), median(pred[mask])]
    as a (3,) float32 CUDA tensor (ofb_median_scale_f32: radix selection, torch.median's lower-median convention),
    stream-ordered, no host synchronisation."""
    _same_shape(pred, gt, mask)
    pred = _lib.require_cuda(pred, "pred")
    gt = _lib.require_cuda(gt, "gt")
    m = (mask > 0).to(torch.uint8).contiguous()
    state = torch.empty(1027, dtype=torch.int32, device=pred.device)
    out = torch.empty(3, dtype=torch.float32, device=pred.device)
    _lib.use_device(pred.device)
    _lib.check(_lib.lib().ofb_median_scale_f32(_lib.ptr(pred), _lib.ptr(gt), _lib.ptr(m), pred.numel(),
                                               _lib.ptr(state), _lib.ptr(out), _lib.stream_of(pred.device)))
    return out


def median_scale(pred, gt, mask):
    """test.py:161-162: median(gt[mask]) / median(pred[mask]) over the whole batch tensor, as a Python float."""
    return median_scale_device(pred, gt, mask)[0].item()


METRIC_NAMES = ("abs_rel", "sq_rel", "rms_sq_lin", "rms_sq_log", "d1", "d2", "d3")


def depth_metrics_partial(pred, gt, mask, scale=1.0):
    """All seven metrics of metrics.py:7-26 as float64 partial sums (9,) from one kernel pass:
    [abs_rel, sq_rel, rms_sq_lin, rms_sq_log sums, n_log, d1, d2, d3 counts, n]."""
    _same_shape(pred, gt, mask)
    pred = _lib.require_cuda(pred, "pred")
    gt = _lib.require_cuda(gt, "gt")
    m = (mask > 0).to(torch.uint8).contiguous()
    out = torch.zeros(9, dtype=torch.float64, device=pred.device)
    _lib.use_device(pred.device)
    if torch.is_tensor(scale):          # device scalar (median_scale_device(...)[0:1]): no host round trip
        _lib.check(_lib.lib().ofb_depth_metrics_partial_ds(_lib.ptr(pred), _lib.ptr(gt), _lib.ptr(m), pred.numel(),
                                                           _lib.ptr(scale), _lib.ptr(out), _lib.stream_of(pred.device)))
    else:
        _lib.check(_lib.lib().ofb_depth_metrics_partial(_lib.ptr(pred), _lib.ptr(gt), _lib.ptr(m), pred.numel(),
                                                        float(scale), _lib.ptr(out), _lib.stream_of(pred.device)))
    return out


def finalize_metrics(partial):
    """(9,) partial sums of ONE batch tensor -> dict of the seven per-batch values test.py:163-169 computes
    (rms_sq_log is a mean over its own valid-pixel count, metrics.py:19-22)."""
    p = partial.tolist()
    n = max(p[8], 1.0)
    return {"abs_rel": p[0] / n, "sq_rel": p[1] / n, "rms_sq_lin": p[2] / n, "rms_sq_log": p[3] / max(p[4], 1.0),
            "d1": p[5] / n, "d2": p[6] / n, "d3": p[7] / n, "n": int(p[8])}


def compute_eval_metrics(pred, gt, mask, use_median_scale=True):
    """test.py:151-170 for one batch tensor: median scaling, then the seven metrics."""
    s = median_scale(pred, gt, mask) if use_median_scale else 1.0
    return finalize_metrics(depth_metrics_partial(pred, gt, mask, s))


def meter_update(partial):
    """One AverageMeter.update(val, N) per metric (test.py:171-177) as an (8,) float64 increment
    [val_abs_rel*N, val_sq_rel*N, val_rms_lin*N, val_rms_log*N, val_d1*N, val_d2*N, val_d3*N, N] with N =
    mask.sum() of the batch.  For six of the metrics val*N is the plain sum; rms_sq_log's val is a mean over ITS
    valid pixels (n_log <= N), so its increment is (sum_log / n_log) * N exactly as the reference weights it."""
    p = partial
    n = p[8]
    log_val = torch.where(p[4] > 0, p[3] / p[4].clamp_min(1.0), torch.zeros_like(p[3]))
    return torch.stack([p[0], p[1], p[2], log_val * n, p[5], p[6], p[7], n])


class DepthMeters:
    """The seven N-weighted running means of test.py:121-148,171-177 kept as sums on the device, so that the
    shards of a multi-GPU evaluation combine with ONE all-reduce(SUM) of 8 doubles (parallel.reduce_sums; NCCL).

    Median scaling (test.py:161-162) is taken over the batch tensor handed to update(), as in the reference, where
    a "batch" is whatever one DataLoader step delivers.  Under parallel.shard_batch every rank sees its own shard
    as the batch, so the scale factor is per shard - the same thing nn.DataParallel users of the reference get
    with a per-GPU batch size; pass the gathered batch if one global median is wanted."""

    def __init__(self, device):
        self.acc = torch.zeros(8, dtype=torch.float64, device=device)

    def update(self, pred, gt, mask, use_median_scale=True):
        """Stream-ordered (median selection, metric reduction and accumulation all stay on the device); returns the
        (9,) partial sums of this batch (finalize_metrics() turns them into per-batch values)."""
        s = median_scale_device(pred, gt, mask)[0:1] if use_median_scale else 1.0
        part = depth_metrics_partial(pred, gt, mask, s)
        self.acc += meter_update(part)
        return part

    def all_reduce(self):
        from . import parallel
        parallel.reduce_sums(self.acc)
        return self

    def result(self):
        a = self.acc.tolist()
        n = max(a[7], 1.0)
        out = {k: a[i] / n for i, k in enumerate(METRIC_NAMES)}
        out["n"] = int(a[7])
        return out


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
