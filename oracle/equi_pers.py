"""Oracle restatement of the two spherical resamplers (TEST INFRASTRUCTURE ONLY).

Follows, step for step and in the same float32 operation order (so that the
integer tap indices come out bit-identical on the same host):

* ``equi2pers``  - /root/reference/equi_pers/equi2pers_v3.py:20-122
* ``pers2equi``  - /root/reference/equi_pers/pers2equi_v3.py:16-198

Differences from the reference that do not change results:
* the dead ERP mask loop (equi2pers_v3.py:49-74) is not evaluated;
* the pers2equi tap table is cached in a process-local dict keyed by the full
  geometry instead of ``./grid/<layer_name>.pth`` on disk keyed by name only
  (pers2equi_v3.py:24-29,155-167) - no stale-cache hazard, no disk traffic.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PI = math.pi
PI_2 = math.pi * 0.5

# rows of tangent patches: (patches per row, row latitude in degrees)
# equi2pers_v3.py:32-47 / pers2equi_v3.py:36-51.  The nrows=3 latitude differs
# between the two files (60 vs 59.6) and both values are kept as they are.
_ROWS = {
    4: ([3, 6, 6, 3], [-67.5, -22.5, 22.5, 67.5]),
    6: ([3, 8, 12, 12, 8, 3], [-75.2, -45.93, -15.72, 15.72, 45.93, 75.2]),
    5: ([3, 6, 8, 6, 3], [-72.2, -36.1, 0, 36.1, 72.2]),
}
_ROWS3_E2P = ([3, 4, 3], [-60, 0, 60])
_ROWS3_P2E = ([3, 4, 3], [-59.6, 0, 59.6])


def _pair(t):
    return t if isinstance(t, tuple) else (t, t)


def patch_centers_deg(nrows, for_pers2equi=False):
    """(N,2) float64 [theta, phi] in degrees; equi2pers_v3.py:51-56,75."""
    if nrows == 3:
        ncols, phis = _ROWS3_P2E if for_pers2equi else _ROWS3_E2P
    else:
        ncols, phis = _ROWS[nrows]  # KeyError for unsupported nrows (reference: UnboundLocalError)
    rows = []
    for i, n_cols in enumerate(ncols):
        for j in np.arange(n_cols):
            theta_interval = 360 / n_cols
            rows.append([j * theta_interval + theta_interval / 2, phis[i]])
    return np.vstack(rows)


def centers_radians(nrows, for_pers2equi=False):
    """Returns (cp[N,1,2] radians float32, center_p[N,2] in [-1,1]).

    equi2pers_v3.py:80-89, pers2equi_v3.py:65-72.
    """
    c = torch.from_numpy(patch_centers_deg(nrows, for_pers2equi)).float()
    c[:, 0] = (c[:, 0]) / 360
    c[:, 1] = (c[:, 1] + 90) / 180
    cp = c * 2 - 1
    center_p = cp.clone()
    cp[:, 0] = cp[:, 0] * PI
    cp[:, 1] = cp[:, 1] * PI_2
    return cp.unsqueeze(1), center_p


def equi2pers_geometry(fov, nrows, patch_size):
    """Input-independent part of equi2pers (equi2pers_v3.py:22-30,86-104).

    Returns dict with
      lon, lat         (N, P*P) float32, un-wrapped radians
      grid             (N, Ph, Pw, 2) float32 [gx, gy] in [-1,1] (lon wrapped)
      center_p         (N, 2)
    """
    height, width = _pair(patch_size)
    fov_h, fov_w = _pair(fov)
    FOV = torch.tensor([fov_w / 360.0, fov_h / 180.0], dtype=torch.float32)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, height), torch.linspace(0, 1, width), indexing="ij")
    screen = torch.stack([xx.flatten(), yy.flatten()], -1)
    cp, center_p = centers_radians(nrows)
    n = cp.shape[0]

    conv = screen * 2 - 1
    conv[:, 0] = conv[:, 0] * PI
    conv[:, 1] = conv[:, 1] * PI_2
    conv = conv * (torch.ones(screen.shape, dtype=torch.float32) * FOV)
    conv = conv.unsqueeze(0).repeat(n, 1, 1)
    x = conv[:, :, 0]
    y = conv[:, :, 1]

    rou = torch.sqrt(x ** 2 + y ** 2)
    c = torch.atan(rou)
    sin_c = torch.sin(c)
    cos_c = torch.cos(c)
    lat = torch.asin(cos_c * torch.sin(cp[:, :, 1]) + (y * sin_c * torch.cos(cp[:, :, 1])) / rou)
    lon = cp[:, :, 0] + torch.atan2(
        x * sin_c, rou * torch.cos(cp[:, :, 1]) * cos_c - y * torch.sin(cp[:, :, 1]) * sin_c)
    lat_new = lat / PI_2
    lon_new = lon / PI
    lon_new[lon_new > 1] -= 2
    lon_new[lon_new < -1] += 2
    grid = torch.stack([lon_new, lat_new], -1).view(n, height, width, 2)
    return {"lon": lon, "lat": lat, "grid": grid, "center_p": center_p}


def uv2xyz(uv):
    """Unit-sphere coordinates; equi2pers_v3.py:13-18 (numpy float32 trig)."""
    xyz = np.zeros((*uv.shape[:-1], 3), dtype=np.float32)
    xyz[..., 0] = np.multiply(np.cos(uv[..., 1]), np.sin(uv[..., 0]))
    xyz[..., 1] = np.multiply(np.cos(uv[..., 1]), np.cos(uv[..., 0]))
    xyz[..., 2] = np.sin(uv[..., 1])
    return xyz


def grid_sample_taps(grid, erp_h, erp_w):
    """Integer neighbour indices F.grid_sample(bilinear, border, align_corners=True)
    uses for ``grid`` (..., 2): returns (x0, y0) int64 and (ix, iy) float32.

    Restates ATen's unnormalise + clip: ix=((gx+1)/2)(W-1),
    clamped to [0, W-1]; x0=floor(ix).
    """
    ix = ((grid[..., 0] + 1) / 2) * (erp_w - 1)
    iy = ((grid[..., 1] + 1) / 2) * (erp_h - 1)
    ix = ix.clamp(0, erp_w - 1)
    iy = iy.clamp(0, erp_h - 1)
    return torch.floor(ix).long(), torch.floor(iy).long(), ix, iy


def equi2pers(erp_img, fov, nrows, patch_size):
    """equi2pers_v3.py:20-122 -> (pers[B,C,Ph,Pw,N], xyz[N,3,Ph,Pw], uv[N,2,Ph,Pw], center_p[N,2])."""
    bs = erp_img.shape[0]
    height, width = _pair(patch_size)
    g = equi2pers_geometry(fov, nrows, patch_size)
    n = g["grid"].shape[0]
    # patches side by side along the width: (Ph, N*Pw, 2)  (equi2pers_v3.py:106-108)
    wide = g["grid"].permute(1, 0, 2, 3).contiguous().view(height, n * width, 2)
    grid = wide.unsqueeze(0).repeat(bs, 1, 1, 1).to(erp_img.device)
    pers = F.grid_sample(erp_img, grid, mode="bilinear", padding_mode="border", align_corners=True)
    pers = F.unfold(pers, kernel_size=(height, width), stride=(height, width))
    pers = pers.reshape(bs, -1, height, width, n)

    grid_tmp = torch.stack([g["lon"], g["lat"]], -1)
    xyz = uv2xyz(grid_tmp)
    xyz = xyz.reshape(n, height, width, 3).transpose(0, 3, 1, 2)
    xyz = torch.from_numpy(xyz).to(pers.device).contiguous()
    # raw reshape of the wide grid, exactly as the reference does (:120-121)
    uv = grid[0, ...].reshape(height, width, n, 2).permute(2, 3, 0, 1).contiguous()
    return pers, xyz, uv, g["center_p"]


def pers2equi_table(fov, nrows, patch_size, erp_size):
    """Tap table of pers2equi (pers2equi_v3.py:30-153).

    Returns dict x0,y0,x1,y1,mask: (N,He,We) int64; w_list: (N,He,We,4) float32.
    """
    height, width = _pair(patch_size)
    fov_h, fov_w = _pair(fov)
    erp_h, erp_w = _pair(erp_size)
    FOV = torch.tensor([fov_w / 360.0, fov_h / 180.0], dtype=torch.float32)
    cp, _ = centers_radians(nrows, for_pers2equi=True)
    n_patch = cp.shape[0]

    lat_grid, lon_grid = torch.meshgrid(
        torch.linspace(-PI_2, PI_2, erp_h), torch.linspace(-PI, PI, erp_w), indexing="ij")
    lon_grid = lon_grid.float().reshape(1, -1)
    lat_grid = lat_grid.float().reshape(1, -1)
    cos_c = torch.sin(cp[..., 1]) * torch.sin(lat_grid) + \
        torch.cos(cp[..., 1]) * torch.cos(lat_grid) * torch.cos(lon_grid - cp[..., 0])
    new_x = (torch.cos(lat_grid) * torch.sin(lon_grid - cp[..., 0])) / cos_c
    new_y = (torch.cos(cp[..., 1]) * torch.sin(lat_grid)
             - torch.sin(cp[..., 1]) * torch.cos(lat_grid) * torch.cos(lon_grid - cp[..., 0])) / cos_c
    new_x = new_x / FOV[0] / PI
    new_y = new_y / FOV[1] / PI_2
    front = torch.where(cos_c.reshape(n_patch, erp_h, erp_w) > 0, 1, 0)

    xp = ((new_x + 1) * 0.5 * height).reshape(n_patch, erp_h, erp_w)
    yp = ((new_y + 1) * 0.5 * width).reshape(n_patch, erp_h, erp_w)
    mask = torch.where((xp < width) & (xp > 0) & (yp < height) & (yp > 0), 1, 0)
    mask *= front

    x0 = torch.floor(xp).type(torch.int64)
    x1 = x0 + 1
    y0 = torch.floor(yp).type(torch.int64)
    y1 = y0 + 1
    x0 = torch.clamp(x0, 0, width - 1)
    x1 = torch.clamp(x1, 0, width - 1)
    y0 = torch.clamp(y0, 0, height - 1)
    y1 = torch.clamp(y1, 0, height - 1)

    x0f, x1f = x0.type(torch.float32), x1.type(torch.float32)
    y0f, y1f = y0.type(torch.float32), y1.type(torch.float32)
    w_list = torch.zeros((n_patch, erp_h, erp_w, 4), dtype=torch.float32)
    w_list[..., 0] = ((x1f - xp) * (y1f - yp)) * mask
    w_list[..., 1] = ((x1f - xp) * (yp - y0f)) * mask
    w_list[..., 2] = ((xp - x0f) * (y1f - yp)) * mask
    w_list[..., 3] = ((xp - x0f) * (yp - y0f)) * mask
    return {"x0": x0, "y0": y0, "x1": x1, "y1": y1, "w_list": w_list, "mask": mask}


def normalized_weights(w_list):
    """Threshold + L1 normalisation over all N*4 taps of a pixel
    (pers2equi_v3.py:189-192).  (N,He,We,4) -> (He,We,N,4)."""
    n_patch, erp_h, erp_w, _ = w_list.shape
    w = w_list.permute(1, 2, 0, 3).flatten(2).clone()
    w *= torch.gt(w, 1e-5).type(torch.float32)
    return F.normalize(w, p=1, dim=-1).reshape(erp_h, erp_w, n_patch, 4)


_TABLES = {}


def _cached_table(fov, nrows, patch_size, erp_size):
    key = (_pair(fov), nrows, _pair(patch_size), _pair(erp_size))
    if key not in _TABLES:
        _TABLES.clear()  # a table is 0.5-3 GB; keep one
        _TABLES[key] = pers2equi_table(fov, nrows, patch_size, erp_size)
    return _TABLES[key]


def pers2equi(pers_img, fov, nrows, patch_size, erp_size, layer_name=None):
    """pers2equi_v3.py:16-198: (B,C,Ph,Pw,N) -> (B,C,He,We).  ``layer_name`` only
    named the reference's disk cache and is ignored."""
    n_patch = pers_img.shape[-1]
    t = _cached_table(fov, nrows, patch_size, erp_size)
    x0, y0, x1, y1 = t["x0"], t["y0"], t["x1"], t["y1"]
    mask = t["mask"].to(pers_img.device)
    z = torch.arange(n_patch).reshape(n_patch, 1, 1)
    taps = [pers_img[:, :, y0, x0, z], pers_img[:, :, y1, x0, z],
            pers_img[:, :, y0, x1, z], pers_img[:, :, y1, x1, z]]
    taps = [(t_ * mask.expand_as(t_)).permute(0, 1, 3, 4, 2) for t_ in taps]
    w = normalized_weights(t["w_list"].to(pers_img.device)).unsqueeze(0).unsqueeze(0)
    out = taps[0] * w[..., 0] + taps[1] * w[..., 1] + taps[2] * w[..., 2] + taps[3] * w[..., 3]
    return out.sum(-1)
