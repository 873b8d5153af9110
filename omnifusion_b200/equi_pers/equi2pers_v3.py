"""ERP -> tangent patches.  Same signature and results as the reference's
equi_pers/equi2pers_v3.py:20-122, computed by the ``ofb_equi2pers_f32`` CUDA kernel; differentiable with respect to
``erp_img`` (``ofb_equi2pers_backward_f32``), like the reference's F.grid_sample."""
import torch

from .. import _lib, tables


def _forward(erp, grid, n, ph, pw):
    bs, ch, he, we = erp.shape
    pers = torch.empty((bs, ch, ph, pw, n), dtype=torch.float32, device=erp.device)
    _lib.use_device(erp.device)
    _lib.check(_lib.lib().ofb_equi2pers_f32(
        _lib.ptr(erp), bs, ch, he, we, _lib.ptr(grid), n, ph, pw,
        _lib.ptr(pers), _lib.LAYOUT_REF, _lib.stream_of(erp.device)))
    return pers


class _Equi2Pers(torch.autograd.Function):
    @staticmethod
    def forward(ctx, erp, grid, n, ph, pw):
        ctx.save_for_backward(grid)
        ctx.meta = (tuple(erp.shape), n, ph, pw)
        return _forward(erp, grid, n, ph, pw)

    @staticmethod
    def backward(ctx, grad_pers):
        (grid,) = ctx.saved_tensors
        (bs, ch, he, we), n, ph, pw = ctx.meta
        g = grad_pers.contiguous().float()
        grad_erp = torch.zeros((bs, ch, he, we), dtype=torch.float32, device=g.device)
        _lib.use_device(g.device)
        _lib.check(_lib.lib().ofb_equi2pers_backward_f32(
            _lib.ptr(g), bs, ch, he, we, _lib.ptr(grid), n, ph, pw, _lib.ptr(grad_erp), _lib.stream_of(g.device)))
        return grad_erp, None, None, None, None


def equi2pers(erp_img, fov, nrows, patch_size):
    """erp_img (B,C,He,We) float32 CUDA -> (pers (B,C,Ph,Pw,N), xyz (N,3,Ph,Pw), uv (N,2,Ph,Pw),
    center_p (N,2) on the CPU), exactly as the reference returns them."""
    erp = _lib.require_cuda(erp_img, "erp_img", allow_grad=True)
    if erp.dim() != 4:
        raise ValueError(f"erp_img must be (B,C,He,We), got {tuple(erp.shape)}")
    ph, pw = tables.pair(patch_size)
    geo = tables.device_patch_geometry(fov, nrows, (ph, pw), erp.device)
    n = geo["grid"].shape[0]
    if torch.is_grad_enabled() and erp.requires_grad:
        pers = _Equi2Pers.apply(erp, geo["grid"], n, ph, pw)
    else:
        pers = _forward(erp, geo["grid"], n, ph, pw)
    return pers, geo["xyz"], geo["uv"], geo["center_p"]
