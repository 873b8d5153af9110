"""``supervision/direct.py`` of the reference (lines 3-27) on the GPU: the reverse-Huber (BerHu) depth loss used by
train_erp_depth_iterative.py:271 / train_erp_depth.py:267 and the masked L1 loss, same signatures and results,
computed by the ``ofb_depth_loss_f32`` CUDA kernels (csrc/loss.cu) and differentiable with respect to ``pred``
(``ofb_depth_loss_backward_f32``).  The BerHu threshold ``c = max|gt - pred| / 5`` is taken over the whole unmasked
batch and is a constant for the gradient, exactly like the reference's ``.item()``."""
import torch

from .. import _lib


def _run(pred, gt, maskf, weights, mode):
    bs = pred.shape[0]
    per = pred.numel() // bs
    dev = pred.device
    _lib.use_device(dev)
    L = _lib.lib()
    work = torch.empty(L.ofb_loss_work_bytes(bs), dtype=torch.uint8, device=dev)
    stats = torch.empty(1 + bs, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    _lib.check(L.ofb_depth_loss_f32(_lib.ptr(pred), _lib.ptr(gt), _lib.ptr(maskf), _lib.ptr(weights) if weights is not None else None,
                                    bs, per, mode, _lib.ptr(work), _lib.ptr(stats), _lib.ptr(loss), _lib.stream_of(dev)))
    return loss, stats


class _DepthLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, maskf, weights, mode):
        loss, stats = _run(pred, gt, maskf, weights, mode)
        ctx.save_for_backward(pred, gt, maskf, weights if weights is not None else pred.new_empty(0), stats)
        ctx.mode = mode
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        pred, gt, maskf, weights, stats = ctx.saved_tensors
        bs = pred.shape[0]
        g = grad_out.contiguous().float()
        grad = torch.empty_like(pred)
        _lib.use_device(pred.device)
        _lib.check(_lib.lib().ofb_depth_loss_backward_f32(
            _lib.ptr(pred), _lib.ptr(gt), _lib.ptr(maskf), _lib.ptr(weights) if weights.numel() else None, bs,
            pred.numel() // bs, ctx.mode, _lib.ptr(stats), _lib.ptr(g), _lib.ptr(grad), _lib.stream_of(pred.device)))
        return grad, None, None, None, None


def _prepare(pred, gt, mask, weights):
    pred = _lib.require_cuda(pred, "pred", allow_grad=True)
    gt = _lib.require_cuda(gt, "gt")
    if gt.shape != pred.shape or mask.shape != pred.shape or (weights is not None and weights.shape != pred.shape):
        raise ValueError(f"pred {tuple(pred.shape)}, gt {tuple(gt.shape)}, mask {tuple(mask.shape)}"
                         + (f", weights {tuple(weights.shape)}" if weights is not None else "") + " must have the same shape")
    if not mask.is_cuda:
        raise _lib.OfbError(f"mask must be a CUDA tensor (got {mask.device})")
    maskf = mask.float().contiguous()                       # the reference multiplies by mask.float()
    if weights is not None:
        weights = _lib.require_cuda(weights, "weights")
    return pred, gt, maskf, weights


def _apply(pred, gt, maskf, weights, mode):
    if torch.is_grad_enabled() and pred.requires_grad:
        return _DepthLoss.apply(pred, gt, maskf, weights, mode)
    return _run(pred, gt, maskf, weights, mode)[0]


def calculate_berhu_loss(pred, gt, mask, weights):
    """supervision/direct.py:3-18.  pred, gt, weights (B, ...) float32 CUDA, mask (B, ...) bool / uint8 / float."""
    return _apply(*_prepare(pred, gt, mask, weights), 1)


def calculate_l1_loss(pred, gt, mask):
    """supervision/direct.py:20-27."""
    return _apply(*_prepare(pred, gt, mask, None), 0)
