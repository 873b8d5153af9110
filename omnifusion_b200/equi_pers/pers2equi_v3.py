"""Tangent patches -> ERP with L1-normalised overlap blending.  Same signature and
results as the reference's equi_pers/pers2equi_v3.py:16-198, computed by the
``ofb_pers2equi_f32`` CUDA kernel from a device-resident CSR table (no ./grid cache); differentiable with respect to
``pers_img`` (``ofb_pers2equi_backward_f32``), like the reference's gather / weighted sum."""
import torch

from .. import _lib, tables


def _forward(pers, tab, he, we):
    bs, ch, ph, pw, n = pers.shape
    out = torch.empty((bs, ch, he, we), dtype=torch.float32, device=pers.device)
    _lib.use_device(pers.device)
    _lib.check(_lib.lib().ofb_pers2equi_f32(
        _lib.ptr(pers), bs, ch, n, ph, pw, _lib.LAYOUT_REF,
        _lib.ptr(tab["rowptr"]), _lib.ptr(tab["idx"]), _lib.ptr(tab["w"]), he, we,
        _lib.ptr(out), _lib.stream_of(pers.device)))
    return out


class _Pers2Equi(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pers, rowptr, idx, w, he, we):
        ctx.save_for_backward(rowptr, idx, w)
        ctx.meta = (tuple(pers.shape), he, we)
        return _forward(pers, {"rowptr": rowptr, "idx": idx, "w": w}, he, we)

    @staticmethod
    def backward(ctx, grad_erp):
        rowptr, idx, w = ctx.saved_tensors
        (bs, ch, ph, pw, n), he, we = ctx.meta
        g = grad_erp.contiguous().float()
        grad_pers = torch.zeros((bs, ch, ph, pw, n), dtype=torch.float32, device=g.device)
        _lib.use_device(g.device)
        _lib.check(_lib.lib().ofb_pers2equi_backward_f32(
            _lib.ptr(g), bs, ch, n, ph, pw, _lib.ptr(rowptr), _lib.ptr(idx), _lib.ptr(w), he, we,
            _lib.ptr(grad_pers), _lib.stream_of(g.device)))
        return grad_pers, None, None, None, None, None


def pers2equi(pers_img, fov, nrows, patch_size, erp_size, layer_name=None):
    """pers_img (B,C,Ph,Pw,N) float32 CUDA -> (B,C,He,We).  ``layer_name`` only named the
    reference's on-disk table cache (pers2equi_v3.py:27) and is accepted and ignored."""
    pers = _lib.require_cuda(pers_img, "pers_img", allow_grad=True)
    if pers.dim() != 5:
        raise ValueError(f"pers_img must be (B,C,Ph,Pw,N), got {tuple(pers.shape)}")
    bs, ch, ph, pw, n = pers.shape
    if (ph, pw) != tables.pair(patch_size):
        raise ValueError(f"patch_size {patch_size} does not match pers_img {tuple(pers.shape)}")
    he, we = tables.pair(erp_size)
    tab = tables.device_blend_table(fov, nrows, (ph, pw), (he, we), pers.device)
    if tab["n_patch"] != n:
        raise ValueError(f"pers_img has {n} patches, nrows={nrows} has {tab['n_patch']}")
    if torch.is_grad_enabled() and pers.requires_grad:
        return _Pers2Equi.apply(pers, tab["rowptr"], tab["idx"], tab["w"], he, we)
    return _forward(pers, tab, he, we)
