"""Input-independent geometry tables, built once on the host and cached on the device.

The reference recomputes these on every call (equi_pers/equi2pers_v3.py:29-104 on the
host; equi_pers/pers2equi_v3.py:109-152 cached to ./grid/<name>.pth and re-loaded from
disk per call).  Transcendental functions cannot be made bit-identical between the CPU
(Sleef) and the GPU (libdevice), and the integer tap indices must match the reference
exactly, so the float32 tables are evaluated on the host with the same torch/numpy
operation order as the reference and only the data-dependent part runs on the GPU.

* ``patch_geometry``  : gnomonic sampling grid + xyz + uv + centres for equi2pers.
* ``blend_table``     : pers2equi's dense (N,He,We) tap table compacted to CSR over ERP
                        pixels, with the thresholded, L1-normalised weights
                        (pers2equi_v3.py:189-192) folded in.
"""
import math
import threading

import numpy as np
import torch
import torch.nn.functional as F

_PI, _PI_2 = math.pi, math.pi * 0.5

# (patches per row, row latitude in degrees): equi2pers_v3.py:32-47, pers2equi_v3.py:36-51
ROW_LAYOUT = {
    3: ((3, 4, 3), (-60, 0, 60)),
    4: ((3, 6, 6, 3), (-67.5, -22.5, 22.5, 67.5)),
    5: ((3, 6, 8, 6, 3), (-72.2, -36.1, 0, 36.1, 72.2)),
    6: ((3, 8, 12, 12, 8, 3), (-75.2, -45.93, -15.72, 15.72, 45.93, 75.2)),
}
# pers2equi_v3.py:47 uses 59.6 (not 60) degrees for nrows == 3
ROW_LAYOUT_BLEND3 = ((3, 4, 3), (-59.6, 0, 59.6))
NUM_PATCHES = {k: sum(v[0]) for k, v in ROW_LAYOUT.items()}


def pair(t):
    return tuple(t) if isinstance(t, (tuple, list)) else (t, t)


def _centers(nrows, blend=False):
    """Patch centres: radians (N,1,2) float32 and the [-1,1] pair (N,2) returned to callers."""
    if nrows not in ROW_LAYOUT:
        # the reference fails with UnboundLocalError at equi2pers_v3.py:49 for other values
        raise ValueError(f"nrows must be one of {sorted(ROW_LAYOUT)}, got {nrows}")
    ncols, phis = ROW_LAYOUT_BLEND3 if (blend and nrows == 3) else ROW_LAYOUT[nrows]
    deg = []
    for n_cols, phi in zip(ncols, phis):
        step = 360 / n_cols
        for j in np.arange(n_cols):
            deg.append([j * step + step / 2, phi])
    c = torch.from_numpy(np.vstack(deg)).float()
    c[:, 0] = c[:, 0] / 360
    c[:, 1] = (c[:, 1] + 90) / 180
    unit = c * 2 - 1
    rad = unit.clone()
    rad[:, 0] = rad[:, 0] * _PI
    rad[:, 1] = rad[:, 1] * _PI_2
    return rad.unsqueeze(1), unit


def patch_geometry(fov, nrows, patch_size):
    """Host tensors for equi2pers: grid (N,Ph,Pw,2), xyz (N,3,Ph,Pw), uv (N,2,Ph,Pw), center_p (N,2)."""
    ph, pw = pair(patch_size)
    fov_h, fov_w = pair(fov)
    scale = torch.tensor([fov_w / 360.0, fov_h / 180.0], dtype=torch.float32)
    vv, uu = torch.meshgrid(torch.linspace(0, 1, ph), torch.linspace(0, 1, pw), indexing="ij")
    screen = torch.stack([uu.flatten(), vv.flatten()], -1)
    rad, unit = _centers(nrows)
    n = rad.shape[0]
    lam0, phi0 = rad[:, :, 0], rad[:, :, 1]

    plane = screen * 2 - 1
    plane[:, 0] = plane[:, 0] * _PI
    plane[:, 1] = plane[:, 1] * _PI_2
    plane = plane * (torch.ones(screen.shape, dtype=torch.float32) * scale)
    plane = plane.unsqueeze(0).repeat(n, 1, 1)
    x, y = plane[:, :, 0], plane[:, :, 1]

    # inverse gnomonic projection, operation order of equi2pers_v3.py:95-100
    rho = torch.sqrt(x ** 2 + y ** 2)
    ang = torch.atan(rho)
    s, c = torch.sin(ang), torch.cos(ang)
    lat = torch.asin(c * torch.sin(phi0) + (y * s * torch.cos(phi0)) / rho)
    lon = lam0 + torch.atan2(x * s, rho * torch.cos(phi0) * c - y * torch.sin(phi0) * s)
    gy = lat / _PI_2
    gx = lon / _PI
    gx[gx > 1] -= 2
    gx[gx < -1] += 2
    grid = torch.stack([gx, gy], -1).view(n, ph, pw, 2).contiguous()

    # unit-sphere coordinates with numpy float32 trig on the un-wrapped angles (:13-18,115-118)
    ll = torch.stack([lon, lat], -1)
    xyz = np.zeros((n, ph * pw, 3), dtype=np.float32)
    xyz[..., 0] = np.multiply(np.cos(ll[..., 1]), np.sin(ll[..., 0]))
    xyz[..., 1] = np.multiply(np.cos(ll[..., 1]), np.cos(ll[..., 0]))
    xyz[..., 2] = np.sin(ll[..., 1])
    xyz = torch.from_numpy(xyz.reshape(n, ph, pw, 3).transpose(0, 3, 1, 2)).contiguous()

    # `uv` is the reference's raw reshape of the patches-side-by-side grid (:106-108,120-121)
    wide = grid.permute(1, 0, 2, 3).reshape(ph, n * pw, 2)
    uv = wide.reshape(ph, pw, n, 2).permute(2, 3, 0, 1).contiguous()
    return {"grid": grid, "xyz": xyz, "uv": uv, "center_p": unit}


def blend_table(fov, nrows, patch_size, erp_size, rows_per_chunk=64):
    """CSR blend table for pers2equi.

    Returns host tensors rowptr (He*We+1,) int32, idx (nnz,) int32 (bit pattern of the packed
    uint32 n<<24 | y0<<16 | x0<<8 | dy<<1 | dx), w (nnz,4) float32, plus ``dense_nnz`` =
    number of (patch, pixel) pairs with mask == 1 in the reference's table.
    """
    ph, pw = pair(patch_size)
    fov_h, fov_w = pair(fov)
    he, we = pair(erp_size)
    if ph > 256 or pw > 256:
        raise ValueError("blend_table packs tap coordinates in 8 bits: patch size must be <= 256")
    scale = torch.tensor([fov_w / 360.0, fov_h / 180.0], dtype=torch.float32)
    rad, _ = _centers(nrows, blend=True)
    lam0, phi0 = rad[..., 0], rad[..., 1]          # (N,1)
    n = rad.shape[0]
    lat_all, lon_all = torch.meshgrid(torch.linspace(-_PI_2, _PI_2, he), torch.linspace(-_PI, _PI, we),
                                      indexing="ij")
    counts, idx_parts, w_parts = [], [], []
    dense_nnz = 0
    for r0 in range(0, he, rows_per_chunk):
        r1 = min(he, r0 + rows_per_chunk)
        rows = r1 - r0
        lon = lon_all[r0:r1].float().reshape(1, -1)
        lat = lat_all[r0:r1].float().reshape(1, -1)
        # forward gnomonic projection, operation order of pers2equi_v3.py:112-116
        cosc = torch.sin(phi0) * torch.sin(lat) + torch.cos(phi0) * torch.cos(lat) * torch.cos(lon - lam0)
        px = (torch.cos(lat) * torch.sin(lon - lam0)) / cosc
        py = (torch.cos(phi0) * torch.sin(lat) - torch.sin(phi0) * torch.cos(lat) * torch.cos(lon - lam0)) / cosc
        px = px / scale[0] / _PI
        py = py / scale[1] / _PI_2
        front = torch.where(cosc.reshape(n, rows, we) > 0, 1, 0)
        xp = ((px + 1) * 0.5 * ph).reshape(n, rows, we)      # sic: height scales x (:122-123)
        yp = ((py + 1) * 0.5 * pw).reshape(n, rows, we)
        mask = torch.where((xp < pw) & (xp > 0) & (yp < ph) & (yp > 0), 1, 0)
        mask *= front
        dense_nnz += int(mask.sum())
        x0 = torch.floor(xp).type(torch.int64)
        y0 = torch.floor(yp).type(torch.int64)
        x1 = torch.clamp(x0 + 1, 0, pw - 1)
        y1 = torch.clamp(y0 + 1, 0, ph - 1)
        x0 = torch.clamp(x0, 0, pw - 1)
        y0 = torch.clamp(y0, 0, ph - 1)
        x0f, x1f, y0f, y1f = (t.type(torch.float32) for t in (x0, x1, y0, y1))
        wl = torch.zeros((n, rows, we, 4), dtype=torch.float32)
        wl[..., 0] = ((x1f - xp) * (y1f - yp)) * mask
        wl[..., 1] = ((x1f - xp) * (yp - y0f)) * mask
        wl[..., 2] = ((xp - x0f) * (y1f - yp)) * mask
        wl[..., 3] = ((xp - x0f) * (yp - y0f)) * mask
        # threshold + L1 normalisation over the N*4 taps of each pixel (:189-192)
        wn = wl.permute(1, 2, 0, 3).flatten(2)
        wn = wn * torch.gt(wn, 1e-5).type(torch.float32)
        wn = F.normalize(wn, p=1, dim=-1).reshape(rows * we, n, 4)
        keep = (wn != 0).any(-1)                              # (pix, N)
        counts.append(keep.sum(1).to(torch.int32))
        pix, pn = torch.nonzero(keep, as_tuple=True)          # pixel-major, patch ascending
        sel = lambda t: t.permute(1, 2, 0).reshape(rows * we, n)[pix, pn]
        sx0, sy0, sx1, sy1 = sel(x0), sel(y0), sel(x1), sel(y1)
        packed = (pn << 24) | (sy0 << 16) | (sx0 << 8) | ((sy1 - sy0) << 1) | (sx1 - sx0)
        idx_parts.append(packed.to(torch.int64))
        w_parts.append(wn[pix, pn])
    counts = torch.cat(counts)
    rowptr = torch.zeros(he * we + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts.to(torch.int64), 0)
    idx = torch.cat(idx_parts)
    # store the uint32 bit pattern in an int32 tensor
    idx = torch.where(idx >= 2 ** 31, idx - 2 ** 32, idx).to(torch.int32)
    return {"rowptr": rowptr.to(torch.int32), "idx": idx.contiguous(),
            "w": torch.cat(w_parts).contiguous(), "dense_nnz": dense_nnz, "n_patch": n}


def unpack_idx(idx):
    """Decode packed CSR indices -> (n, y0, x0, y1, x1) int64 tensors (used by tests)."""
    v = idx.to(torch.int64) & 0xFFFFFFFF
    n, y0, x0 = v >> 24, (v >> 16) & 255, (v >> 8) & 255
    return n, y0, x0, y0 + ((v >> 1) & 1), x0 + (v & 1)


# ----------------------------------------------------------------- device caches
_lock = threading.Lock()
_patch_cache = {}
_blend_cache = {}
_BLEND_CACHE_MAX = 4


def device_patch_geometry(fov, nrows, patch_size, device):
    key = (pair(fov), nrows, pair(patch_size), str(device))
    with _lock:
        g = _patch_cache.get(key)
        if g is None:
            host = patch_geometry(fov, nrows, patch_size)
            g = {k: (v.to(device) if k != "center_p" else v) for k, v in host.items()}
            _patch_cache[key] = g
    return g


def device_blend_table(fov, nrows, patch_size, erp_size, device):
    key = (pair(fov), nrows, pair(patch_size), pair(erp_size), str(device))
    with _lock:
        t = _blend_cache.get(key)
        if t is None:
            host = blend_table(fov, nrows, patch_size, erp_size)
            t = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in host.items()}
            if len(_blend_cache) >= _BLEND_CACHE_MAX:
                _blend_cache.pop(next(iter(_blend_cache)))
            _blend_cache[key] = t
    return t
