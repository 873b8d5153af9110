"""ERP -> tangent patches.  Same signature and results as the reference's
equi_pers/equi2pers_v3.py:20-122, computed by the ``ofb_equi2pers_f32`` CUDA kernel."""
import torch

from .. import _lib, tables


def equi2pers(erp_img, fov, nrows, patch_size):
    """erp_img (B,C,He,We) float32 CUDA -> (pers (B,C,Ph,Pw,N), xyz (N,3,Ph,Pw), uv (N,2,Ph,Pw),
    center_p (N,2) on the CPU), exactly as the reference returns them."""
    erp = _lib.require_cuda(erp_img, "erp_img")
    if erp.dim() != 4:
        raise ValueError(f"erp_img must be (B,C,He,We), got {tuple(erp.shape)}")
    bs, ch, he, we = erp.shape
    ph, pw = tables.pair(patch_size)
    geo = tables.device_patch_geometry(fov, nrows, (ph, pw), erp.device)
    n = geo["grid"].shape[0]
    pers = torch.empty((bs, ch, ph, pw, n), dtype=torch.float32, device=erp.device)
    _lib.use_device(erp.device)
    _lib.check(_lib.lib().ofb_equi2pers_f32(
        _lib.ptr(erp), bs, ch, he, we, _lib.ptr(geo["grid"]), n, ph, pw,
        _lib.ptr(pers), _lib.LAYOUT_REF, _lib.stream_of(erp.device)))
    return pers, geo["xyz"], geo["uv"], geo["center_p"]
