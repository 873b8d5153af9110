"""Point cloud of an ERP depth map, the compute part of the reference's export (test.py:205-218 with
util.py:159-174 coords2uv / uv2xyz): one unit ray per ERP pixel (host numpy table, the reference's exact op
order) scaled by the predicted depth by a libofb kernel.  Writing the PLY file is left to the caller (ply.py)."""
import numpy as np
import torch

from . import _lib


def erp_rays(h, w):
    """(h*w, 3) float32 unit rays, pixel order = row-major over (y, x) like np.meshgrid(range(w), range(h)).
    test.py:210-215: coords = meshgrid + 1; util.py:159-174: uv then xyz."""
    coords = np.stack(np.meshgrid(range(w), range(h)), -1)
    coords = np.reshape(coords, [-1, 2])
    coords = coords + 1
    uv = np.zeros_like(coords, dtype=np.float32)
    middle_x = w / 2 + 0.5
    middle_y = h / 2 + 0.5
    uv[..., 0] = (coords[..., 0] - middle_x) / w * 2 * np.pi
    uv[..., 1] = -(coords[..., 1] - middle_y) / h * np.pi
    xyz = np.zeros((uv.shape[0], 3), dtype=np.float32)
    xyz[:, 0] = np.multiply(np.cos(uv[:, 1]), np.sin(uv[:, 0]))
    xyz[:, 1] = np.multiply(np.cos(uv[:, 1]), np.cos(uv[:, 0]))
    xyz[:, 2] = np.sin(uv[:, 1])
    return xyz


_rays_cache = {}


def depth_to_points(depth, max_depth=None):
    """depth (B,1,H,W) float32 CUDA -> (B, H*W, 3) points = ray * depth (test.py:216-218) in one kernel
    (ofb_depth_to_points_f32).  max_depth: the reference zeroes predictions above 8 m before visualising them
    (test.py:208)."""
    depth = _lib.require_cuda(depth, "depth")
    b, c, h, w = depth.shape
    if c != 1:
        raise ValueError(f"depth must be (B,1,H,W), got {tuple(depth.shape)}")
    key = (h, w, str(depth.device))
    rays = _rays_cache.get(key)
    if rays is None:
        rays = torch.from_numpy(erp_rays(h, w)).to(depth.device)
        _rays_cache.clear()
        _rays_cache[key] = rays
    pts = torch.empty(b, h * w, 3, dtype=torch.float32, device=depth.device)
    _lib.use_device(depth.device)
    _lib.check(_lib.lib().ofb_depth_to_points_f32(_lib.ptr(depth), _lib.ptr(rays), b, h, w,
                                                  float(max_depth) if max_depth is not None else 0.0,
                                                  _lib.ptr(pts), _lib.stream_of(depth.device)))
    return pts
