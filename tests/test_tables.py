"""Host-built geometry tables of the product vs the oracle's dense tables (bit-exact)."""
import numpy as np
import pytest
import torch

from omnifusion_b200 import tables
from oracle import equi_pers as oe

FOV = (80, 80)


@pytest.mark.parametrize("nrows,P", [(3, 16), (4, 32), (5, 16), (6, 128)])
def test_patch_geometry_matches_oracle(nrows, P):
    g = tables.patch_geometry(FOV, nrows, (P, P))
    o = oe.equi2pers_geometry(FOV, nrows, (P, P))
    assert torch.equal(g["grid"], o["grid"])
    assert torch.equal(g["center_p"], o["center_p"])
    img = torch.rand(1, 1, 8, 16)
    _, xyz, uv, _ = oe.equi2pers(img, FOV, nrows, (P, P))
    assert torch.equal(g["xyz"], xyz)
    assert torch.equal(g["uv"], uv)
    assert g["grid"].shape[0] == tables.NUM_PATCHES[nrows]


def test_unsupported_nrows_raises():
    with pytest.raises(ValueError):
        tables.patch_geometry(FOV, 7, 16)


@pytest.mark.parametrize("nrows,erp,P", [(3, (32, 64), 16), (4, (64, 128), 32), (5, (48, 96), 16), (6, (32, 64), 128)])
def test_blend_table_matches_oracle(nrows, erp, P):
    t = tables.blend_table(FOV, nrows, (P, P), erp, rows_per_chunk=24)
    o = oe.pers2equi_table(FOV, nrows, (P, P), erp)
    wn = oe.normalized_weights(o["w_list"])                      # (He,We,N,4)
    he, we = erp
    n_patch = wn.shape[2]
    keep = (wn != 0).any(-1).reshape(he * we, n_patch)
    assert torch.equal(t["rowptr"][1:].long(), torch.cumsum(keep.sum(1), 0))
    pix, pn = torch.nonzero(keep, as_tuple=True)
    n, y0, x0, y1, x1 = tables.unpack_idx(t["idx"])
    assert torch.equal(n, pn)
    flat = lambda a: a.permute(1, 2, 0).reshape(he * we, n_patch)[pix, pn]
    # integer taps must decode to exactly the reference's clamped int64 indices
    assert torch.equal(x0, flat(o["x0"])) and torch.equal(y0, flat(o["y0"]))
    assert torch.equal(x1, flat(o["x1"])) and torch.equal(y1, flat(o["y1"]))
    assert torch.equal(t["w"], wn.reshape(he * we, n_patch, 4)[pix, pn])
    assert t["dense_nnz"] == int(o["mask"].sum())


def test_blend_table_chunking_invariant():
    a = tables.blend_table(FOV, 4, 16, (40, 80), rows_per_chunk=7)
    b = tables.blend_table(FOV, 4, 16, (40, 80), rows_per_chunk=64)
    for k in ("rowptr", "idx", "w"):
        assert torch.equal(a[k], b[k])


def test_pointcloud_rays_match_reference_golden(golden_dir):
    """omnifusion_b200.pointcloud.erp_rays vs the reference's coords2uv / uv2xyz (util.py:159-174) run on the
    meshgrid of test.py:210-213 (tests/golden/make_golden_pointcloud.py)."""
    import os
    from omnifusion_b200.pointcloud import erp_rays
    g = np.load(os.path.join(golden_dir, "pointcloud_rays.npz"))
    for key in g.files:
        h, w = (int(v) for v in key.split("_")[1].split("x"))
        got = erp_rays(h, w)
        assert got.dtype == np.float32 and np.array_equal(got, g[key]), key
