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


# --------------------------------------------------------------------------------------------------
# The PRODUCT's CSR table at the BASELINE geometries against the committed outputs of the real reference
# (tests/golden/resample_full_*.npz: the reference's dense x0/y0/x1/y1/mask/w_list, strided).
FULL_CASES = [("full_n4", 4, (512, 1024)), ("full_n6", 6, (512, 1024)), ("full_n5_2k", 5, (1024, 2048))]


@pytest.mark.parametrize("tag,nrows,erp", FULL_CASES, ids=[c[0] for c in FULL_CASES])
def test_product_blend_table_decodes_to_reference_fixture_at_baseline_sizes(tag, nrows, erp, golden_dir):
    import os
    z = np.load(os.path.join(golden_dir, f"resample_{tag}.npz"))
    assert int(z["nrows"]) == nrows and tuple(int(v) for v in z["erp"]) == erp and int(z["P"]) == 128
    s = int(z["stride"])
    he, we = erp
    t = tables.blend_table(FOV, nrows, (128, 128), erp)
    n_patch = tables.NUM_PATCHES[nrows]
    # every dense entry with mask == 1 of the reference is a candidate; the CSR keeps those whose thresholded,
    # L1-normalised weight vector is non-zero (pers2equi_v3.py:189-192)
    assert t["dense_nnz"] == int(z["t_sums"][4]), "count of mask == 1 entries over the WHOLE table"
    ys, xs = torch.arange(0, he, s), torch.arange(0, we, s)
    pix = (ys[:, None] * we + xs[None, :]).reshape(-1)                       # strided ERP pixels, row-major
    rp = t["rowptr"].long()
    beg, end = rp[pix], rp[pix + 1]
    cnt = end - beg
    sel = torch.repeat_interleave(beg, cnt) + (torch.arange(int(cnt.sum())) -
                                               torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt))
    which = torch.repeat_interleave(torch.arange(pix.numel()), cnt)          # index into the strided pixel list
    n, y0, x0, y1, x1 = tables.unpack_idx(t["idx"][sel])
    # reference side: normalise the fixture's own weights exactly as the reference does
    w_ref = torch.from_numpy(z["t_w"]).permute(1, 2, 0, 3).reshape(pix.numel(), n_patch * 4)
    wn = w_ref * torch.gt(w_ref, 1e-5).float()
    wn = torch.nn.functional.normalize(wn, p=1, dim=-1).reshape(pix.numel(), n_patch, 4)
    keep = (wn != 0).any(-1)
    rpix, rn = torch.nonzero(keep, as_tuple=True)
    assert torch.equal(which, rpix) and torch.equal(n, rn), "CSR rows hold exactly the reference's contributing patches"
    dense = lambda k: torch.from_numpy(z["t_" + k].astype(np.int64)).permute(1, 2, 0).reshape(pix.numel(), n_patch)[rpix, rn]
    assert (dense("mask") == 1).all()
    for name, got in (("x0", x0), ("y0", y0), ("x1", x1), ("y1", y1)):
        assert torch.equal(got, dense(name)), f"{name}: product CSR taps differ from the reference's table"
    assert torch.equal(t["w"][sel], wn[rpix, rn]), "pre-normalised weights differ from the reference's"
    # ERP pixels where the reference's table holds NaN weights (cos_c == 0 exactly: x / 0 = inf, inf * mask 0 = NaN,
    # pers2equi_v3.py:114,137-140): L1 normalisation makes the whole pixel NaN in the reference, so the CSR rows of
    # exactly those pixels must carry NaN too (two pixels at 1024x2048 / nrows=5, none at 512x1024)
    nan_pix_ref = torch.unique(torch.from_numpy(z["t_w_nan"]).long()[:, 1] * we + torch.from_numpy(z["t_w_nan"]).long()[:, 2]) \
        if "t_w_nan" in z.files and len(z["t_w_nan"]) else torch.zeros(0, dtype=torch.long)
    nan_entries = torch.nonzero(torch.isnan(t["w"]).any(-1)).flatten()
    nan_pix = torch.unique(torch.searchsorted(rp, nan_entries, right=True) - 1)
    assert torch.equal(nan_pix, nan_pix_ref)
    if tag == "full_n5_2k":
        assert nan_pix.numel() == 2
