"""Parity of every libofb CUDA kernel (called through the C ABI) against the CPU oracle /
the torch-CPU fp32 op the reference uses.  Integer index work: bit-exact.  fp32 value work:
tolerances stated per test."""
import pytest
import torch
import torch.nn.functional as F

from omnifusion_b200 import _lib, tables
from oracle import equi_pers as oe

pytestmark = pytest.mark.gpu
FOV = (80, 80)
DEV = "cuda:0"


def ops():
    import ofb_ops
    return ofb_ops


def rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def urand(*shape, seed=0):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed))


def report(name, got, ref, atol, rtol=0.0):
    err = (got - ref).abs()
    bound = atol + rtol * ref.abs()
    worst = (err - bound).max().item()
    print(f"[parity] {name}: max_abs_err={err.max().item():.3e} ref_absmax={ref.abs().max().item():.3e}")
    assert worst <= 0, f"{name}: max abs err {err.max().item():.3e} exceeds atol={atol} rtol={rtol}"


# ------------------------------------------------------------------ equi2pers
@pytest.mark.parametrize("nrows,erp,P,C", [(4, (512, 1024), 128, 3), (4, (512, 1024), 32, 1), (5, (1024, 2048), 128, 3),
                                           (6, (512, 1024), 128, 3), (3, (64, 128), 16, 2), (6, (256, 512), 32, 1)])
def test_equi2pers_taps_and_values(nrows, erp, P, C):
    o = ops()
    geo = tables.patch_geometry(FOV, nrows, (P, P))
    grid = geo["grid"].to(DEV)
    x0, y0 = o.equi2pers_taps(grid, *erp)
    rx0, ry0, _, _ = oe.grid_sample_taps(geo["grid"], *erp)
    assert torch.equal(x0.cpu().long(), rx0), "x0 taps differ from the reference's grid_sample"
    assert torch.equal(y0.cpu().long(), ry0), "y0 taps differ from the reference's grid_sample"
    img = urand(2, C, *erp, seed=nrows)
    ref, _, _, _ = oe.equi2pers(img, FOV, nrows, (P, P))
    got = o.equi2pers(img.to(DEV), grid, _lib.LAYOUT_REF).cpu()
    report("equi2pers ref-layout", got, ref, atol=1e-6)      # inputs in [0,1]
    if C in (1, 3):
        folded = o.equi2pers(img.to(DEV), grid, _lib.LAYOUT_FOLDED).cpu()      # (B*N,P,P,Cp)
        n = grid.shape[0]
        f = folded[..., :C].reshape(2, n, P, P, C).permute(0, 4, 2, 3, 1)
        report("equi2pers folded-layout", f, ref, atol=1e-6)
        if C == 3:
            assert (folded[..., 3] == 0).all()


def test_equi2pers_public_api_matches_oracle():
    from omnifusion_b200.equi_pers.equi2pers_v3 import equi2pers
    img = urand(3, 3, 256, 512, seed=5)
    ref = oe.equi2pers(img, FOV, 4, (64, 64))
    got = equi2pers(img.to(DEV), FOV, 4, (64, 64))
    report("equi2pers api pers", got[0].cpu(), ref[0], atol=1e-6)
    assert torch.equal(got[1].cpu(), ref[1]) and got[1].is_cuda          # xyz
    assert torch.equal(got[2].cpu(), ref[2]) and got[2].is_cuda          # uv
    assert torch.equal(got[3], ref[3]) and not got[3].is_cuda            # center_p stays on the CPU
    assert got[0].shape == (3, 3, 64, 64, 18) and got[0].is_contiguous()
    # int patch_size / int fov are accepted like the reference's pair()
    got2 = equi2pers(img.to(DEV), 80, 4, 64)
    assert torch.equal(got2[0], got[0])


def test_equi2pers_rejects_cpu_and_bad_nrows():
    from omnifusion_b200.equi_pers.equi2pers_v3 import equi2pers
    with pytest.raises(_lib.OfbError):
        equi2pers(torch.rand(1, 3, 64, 128), FOV, 4, 16)
    with pytest.raises(ValueError):
        equi2pers(torch.rand(1, 3, 64, 128, device=DEV), FOV, 7, 16)


# ------------------------------------------------------------------ pers2equi
@pytest.mark.parametrize("nrows,erp,P,B,C", [(4, (512, 1024), 128, 1, 1), (4, (256, 512), 32, 2, 3), (3, (128, 256), 32, 1, 2),
                                             (5, (128, 256), 64, 9, 1), (6, (256, 512), 128, 1, 1)])
def test_pers2equi_matches_oracle(nrows, erp, P, B, C):
    o = ops()
    n = tables.NUM_PATCHES[nrows]
    pers = urand(B, C, P, P, n, seed=nrows + 7)
    ref = oe.pers2equi(pers, FOV, nrows, (P, P), erp)
    tab = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in tables.blend_table(FOV, nrows, (P, P), erp).items()}
    got = o.pers2equi(pers.to(DEV), tab, *erp, _lib.LAYOUT_REF).cpu()
    report("pers2equi ref-layout", got, ref, atol=2e-6)
    folded = pers.permute(0, 4, 2, 3, 1).reshape(B * n, P, P, C).contiguous()
    got2 = o.pers2equi(folded.to(DEV), tab, *erp, _lib.LAYOUT_FOLDED, dims=(B, C, n, P, P)).cpu()
    report("pers2equi folded-layout", got2, ref, atol=2e-6)
    # constant input -> exactly-normalised weights give back the constant wherever a patch covers
    ones = torch.ones(1, 1, P, P, n, device=DEV)
    back = o.pers2equi(ones, tab, *erp, _lib.LAYOUT_REF).cpu()
    covered = (tab["rowptr"][1:] > tab["rowptr"][:-1]).cpu().reshape(1, 1, *erp)
    assert (back[covered] - 1).abs().max() <= 2e-6
    assert (back[~covered] == 0).all()
    if nrows != 3:
        assert covered.all()


def test_pers2equi_public_api_roundtrip():
    from omnifusion_b200.equi_pers.equi2pers_v3 import equi2pers
    from omnifusion_b200.equi_pers.pers2equi_v3 import pers2equi
    img = urand(2, 3, 128, 256, seed=3)
    pers, _, _, _ = equi2pers(img.to(DEV), FOV, 4, (32, 32))
    back = pers2equi(pers, FOV, 4, (32, 32), (128, 256), "any_name")
    ref_p = oe.equi2pers(img, FOV, 4, (32, 32))[0]
    ref = oe.pers2equi(ref_p, FOV, 4, (32, 32), (128, 256))
    report("pers2equi(equi2pers(x))", back.cpu(), ref, atol=3e-6)
    with pytest.raises(ValueError):
        pers2equi(pers, FOV, 5, (32, 32), (128, 256), "x")     # 18 patches but nrows=5 has 26


def test_blend_conf_matches_oracle():
    o = ops()
    nrows, erp, P, B = 4, (128, 256), 32, 5
    n = tables.NUM_PATCHES[nrows]
    pred = urand(B * n, P, P, seed=1) * 3
    conf = urand(B * n, P, P, seed=2)
    tab = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in tables.blend_table(FOV, nrows, (P, P), erp).items()}
    got = o.blend_conf((pred * conf).to(DEV), conf.to(DEV), B, n, tab, *erp).cpu()
    unf = lambda t: t.reshape(B, n, 1, P, P).permute(0, 2, 3, 4, 1)
    W = oe.pers2equi(unf(conf), FOV, nrows, (P, P), erp)
    D = oe.pers2equi(unf(pred * conf), FOV, nrows, (P, P), erp)
    ref = D / (W + 1e-8 * (W <= 1e-8).float())
    report("blend_conf", got, ref, atol=0, rtol=1e-5)


# ------------------------------------------------------------------------ conv
CONV_CASES = [
    # n, h, w, c0, c1, cout, k, stride, pad, residual, act
    (3, 32, 32, 64, 0, 64, 3, 1, 1, True, 1),
    (2, 32, 32, 64, 0, 128, 3, 2, 1, False, 1),
    (2, 32, 32, 64, 0, 128, 1, 2, 0, False, 0),
    (5, 8, 8, 256, 256, 128, 3, 1, 1, False, 1),
    (2, 64, 64, 64, 64, 32, 3, 1, 1, False, 1),
    (1, 128, 128, 32, 0, 32, 3, 1, 1, False, 1),
    (7, 4, 4, 512, 0, 512, 3, 1, 1, True, 1),
    (37, 1, 1, 512, 0, 2048, 1, 1, 0, False, 2),
    (144, 1, 1, 2048, 0, 512, 1, 1, 0, True, 0),
    (50, 4, 4, 512, 0, 32, 1, 1, 0, False, 0),
]


def _conv_ref(x0, x1, w, scale, shift, res, k, stride, pad, act):
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    y = F.conv2d(x, w, None, stride, pad)
    y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if res is not None:
        y = y + res
    if act == 1:
        y = F.relu(y)
    elif act == 2:
        y = F.gelu(y)
    return y


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_simt_matches_torch_cpu(case):
    n, h, w, c0, c1, cout, k, stride, pad, use_res, act = case
    o = ops()
    x0 = rand(n, c0, h, w, seed=1)
    x1 = rand(n, c1, h, w, seed=2) if c1 else None
    wt = rand(cout, c0 + c1, k, k, seed=3, scale=(1.0 / ((c0 + c1) * k * k)) ** 0.5)
    scale = 0.5 + urand(cout, seed=4)
    shift = rand(cout, seed=5, scale=0.1)
    oh = (h + 2 * pad - k) // stride + 1
    res = rand(n, cout, oh, oh if h == w else (w + 2 * pad - k) // stride + 1, seed=6) if use_res else None
    ref = _conv_ref(x0, x1, wt, scale, shift, res, k, stride, pad, act)
    got = o.conv(o.nhwc(x0).to(DEV), o.ohwi(wt).to(DEV), k, stride, pad,
                 in1=o.nhwc(x1).to(DEV) if c1 else None, scale=scale.to(DEV), shift=shift.to(DEV),
                 residual=o.nhwc(res).to(DEV) if use_res else None, act=act, engine=_lib.ENGINE_SIMT)
    # fp32 accumulation in a different order than mkldnn: K up to 4608 terms
    report(f"conv_simt {case}", o.nchw(got.cpu()), ref, atol=2e-5, rtol=2e-5)


def test_stem_matches_torch_cpu():
    o = ops()
    x = urand(3, 3, 128, 128, seed=1)
    wt = rand(64, 3, 7, 7, seed=2, scale=0.1)
    scale, shift = 0.5 + urand(64, seed=3), rand(64, seed=4, scale=0.1)
    ref = F.relu(F.conv2d(x, wt, None, 2, 3) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    x4 = torch.cat([x, torch.full((3, 1, 128, 128), 7.0)], 1)        # 4th channel must be ignored
    w4 = torch.cat([wt, torch.zeros(64, 1, 7, 7)], 1)
    w_khwc_o = w4.permute(2, 3, 1, 0).contiguous()                   # (7,7,4,64)
    got = o.stem(o.nhwc(x4).to(DEV), w_khwc_o.to(DEV), scale.to(DEV), shift.to(DEV))
    report("stem", o.nchw(got.cpu()), ref, atol=1e-5, rtol=1e-5)


def test_maxpool_and_upsample_match_torch_cpu():
    o = ops()
    x = rand(3, 64, 16, 16, seed=1)
    got = o.maxpool(o.nhwc(x).to(DEV))
    assert torch.equal(o.nchw(got.cpu()), F.max_pool2d(x, 3, 2, 1))
    for (c, hw) in [(512, 4), (64, 32), (32, 64)]:
        x = rand(2, c, hw, hw, seed=c)
        ref = F.interpolate(x, size=(2 * hw, 2 * hw), mode="bilinear", align_corners=False)
        got = o.upsample2x(o.nhwc(x).to(DEV))
        report(f"upsample2x c={c}", o.nchw(got.cpu()), ref, atol=1e-6, rtol=1e-6)
    x = rand(4, 512, 4, 4, seed=9)
    bias = rand(4, 512, seed=10)
    ref = F.interpolate(x + bias.view(4, 512, 1, 1), size=(8, 8), mode="bilinear", align_corners=False)
    got = o.upsample2x(o.nhwc(x).to(DEV), bias.to(DEV))
    report("upsample2x + token bias", o.nchw(got.cpu()), ref, atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize("cin,with_depth", [(3, False), (3, True), (5, False)])
def test_point_embed_matches_torch_cpu(cin, with_depth):
    o = ops()
    N, pp, B = 18, 32, 2
    pts = rand(N, cin, pp, pp, seed=1)
    w1, w2 = rand(16, cin, 1, 1, seed=2), rand(64, 16, 1, 1, seed=3, scale=0.3)
    s1, t1, s2, t2 = 0.5 + urand(16, seed=4), rand(16, seed=5, scale=0.1), 0.5 + urand(64, seed=6), rand(64, seed=7, scale=0.1)
    depth = (0.5 + urand(B * N, pp, pp, seed=8)) if with_depth else None
    base = rand(B * N, 64, pp, pp, seed=9)
    xin = pts.unsqueeze(0).expand(B, -1, -1, -1, -1).reshape(B * N, cin, pp, pp)
    if with_depth:
        xin = xin * depth.unsqueeze(1)
    hid = F.relu(F.conv2d(xin, w1) * s1.view(1, -1, 1, 1) + t1.view(1, -1, 1, 1))
    ref = F.relu(F.conv2d(hid, w2) * s2.view(1, -1, 1, 1) + t2.view(1, -1, 1, 1)) + base
    got = o.point_embed(pts.to(DEV), depth.to(DEV) if with_depth else None, B * N, w1.reshape(16, cin).contiguous().to(DEV),
                        s1.to(DEV), t1.to(DEV), w2.reshape(64, 16).contiguous().to(DEV), s2.to(DEV), t2.to(DEV),
                        o.nhwc(base).to(DEV))
    report("point_embed", o.nchw(got.cpu()), ref, atol=1e-5, rtol=1e-5)


def test_token_pack_layernorm_attention_match_torch_cpu():
    o = ops()
    B, N = 3, 18
    down = rand(B * N, 32, 4, 4, seed=1)                 # NCHW as the reference holds it
    pos = rand(1, N, 512, seed=2, scale=0.02)
    ref = down.reshape(B, N, 512) + pos
    got = o.token_pack(o.nhwc(down).to(DEV), pos.to(DEV), N)
    assert torch.equal(got.cpu().reshape(B, N, 512), ref)
    x = rand(B * N, 512, seed=3, scale=3.0) + 1.5
    g, b = 0.5 + urand(512, seed=4), rand(512, seed=5, scale=0.1)
    for eps in (1e-5, 1e-6):
        report("layernorm", o.layernorm(x.to(DEV), g.to(DEV), b.to(DEV), eps).cpu(),
               F.layer_norm(x, (512,), g, b, eps), atol=2e-6, rtol=2e-6)
    for n_tok in (18, 46, 10):
        q, kv = rand(B * n_tok, 512, seed=6), rand(B * n_tok, 1024, seed=7)
        qh = q.reshape(B, n_tok, 4, 128).permute(0, 2, 1, 3)
        kvh = kv.reshape(B, n_tok, 2, 4, 128).permute(2, 0, 3, 1, 4)
        att = ((qh @ kvh[0].transpose(-2, -1)) * 128 ** -0.5).softmax(-1)
        ref = (att @ kvh[1]).transpose(1, 2).reshape(B * n_tok, 512)
        report(f"attention N={n_tok}", o.attention(q.to(DEV), kv.to(DEV), B, n_tok).cpu(), ref, atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("confidence", [False, True])
def test_heads_match_torch_cpu(confidence):
    o = ops()
    x = F.relu(rand(3, 32, 64, 48 + 16, seed=1))
    wp, wc = rand(1, 32, 3, 3, seed=2, scale=0.1), rand(1, 32, 3, 3, seed=3, scale=0.1)
    bp, bc = 0.7, -0.3
    pred = F.relu(F.conv2d(x, wp, torch.tensor([bp]), 1, 1))
    conf = torch.sigmoid(F.conv2d(x, wc, torch.tensor([bc]), 1, 1))
    gp, gc = o.heads(o.nhwc(x).to(DEV), o.ohwi(wp).to(DEV), bp, o.ohwi(wc).to(DEV), bc, confidence)
    if confidence:
        report("heads pred*conf", gp.cpu(), (pred * conf)[:, 0], atol=2e-6, rtol=1e-5)
        report("heads conf", gc.cpu(), conf[:, 0], atol=2e-6, rtol=1e-5)
    else:
        report("heads pred", gp.cpu(), pred[:, 0], atol=2e-6, rtol=1e-5)


def test_absrel_partial_matches_metrics_py():
    o = ops()
    pred = 0.1 + 7.9 * urand(2, 1, 128, 256, seed=1)
    gt = 0.1 + 7.9 * urand(2, 1, 128, 256, seed=2)
    mask = (gt <= 8) & (gt > 0.1) & (urand(2, 1, 128, 256, seed=3) > 0.2)
    ref = ((pred[mask] - gt[mask]).abs() / gt[mask]).mean()
    s = o.absrel(pred.to(DEV), gt.to(DEV), mask.to(DEV)).cpu()
    assert int(s[1]) == int(mask.sum())
    assert abs(s[0] / s[1] - ref.item()) <= 1e-6


def test_depth_metrics_match_oracle_eval():
    from omnifusion_b200 import metrics
    from oracle import model as om
    pred = 0.1 + 7.9 * urand(2, 1, 128, 256, seed=1)
    gt = 0.1 + 7.9 * urand(2, 1, 128, 256, seed=2)
    mask = (gt <= 8) & (gt > 0.1) & (urand(2, 1, 128, 256, seed=3) > 0.2)
    for med in (False, True):
        ref = om.eval_metrics(pred, gt, mask, median_scale=med)
        got = metrics.compute_eval_metrics(pred.to(DEV), gt.to(DEV), mask.to(DEV), use_median_scale=med)
        assert got["n"] == ref["n"]
        for k in metrics.METRIC_NAMES:
            assert abs(got[k] - ref[k]) <= 2e-6 * max(1.0, abs(ref[k])), (k, got[k], ref[k])


def test_rgb_u8_to_input_is_bit_identical_to_the_loader_expression():
    """dataset_loader_stanford.py:52,79: rgb.astype(np.float32) / 255, transpose(2,0,1)."""
    import numpy as np
    from omnifusion_b200.preprocess import rgb_u8_to_input
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(2, 37, 53, 3), dtype=np.uint8)
    img[0, 0, 0] = (0, 255, 254)
    want = np.stack([(im.astype(np.float32) / 255).transpose(2, 0, 1) for im in img])
    got = rgb_u8_to_input(torch.from_numpy(img).to(DEV)).cpu().numpy()
    assert got.dtype == np.float32 and got.shape == (2, 3, 37, 53)
    assert np.array_equal(got, want)
    with pytest.raises(Exception):
        rgb_u8_to_input(torch.from_numpy(img))          # CPU tensor: no CPU path


def test_depth_to_points_matches_the_reference_expression():
    """test.py:208-218: xyz rays (util.py:159-174) * depth, predictions above 8 m zeroed for the export."""
    from omnifusion_b200.pointcloud import depth_to_points, erp_rays
    depth = 0.1 + 9.9 * torch.rand(2, 1, 16, 32, generator=torch.Generator().manual_seed(3))
    rays = torch.from_numpy(erp_rays(16, 32))
    want = rays.unsqueeze(0) * depth.reshape(2, 16 * 32, 1)
    got = depth_to_points(depth.to(DEV))
    assert torch.equal(got.cpu(), want)
    clipped = depth.clone()
    clipped[clipped > 8] = 0
    assert torch.equal(depth_to_points(depth.to(DEV), max_depth=8).cpu(), rays.unsqueeze(0) * clipped.reshape(2, -1, 1))


def test_area_resize_is_bit_identical_to_cv2_inter_area(golden_dir):
    """dataset_loader_stanford.py:92-94: cv2.resize(rgb, (pano_w, pano_h), interpolation=cv2.INTER_AREA) for the
    integer factors the Stanford panoramas need; fixture produced by cv2 (tests/golden/make_golden_resize.py)."""
    import os
    import numpy as np
    from omnifusion_b200.preprocess import area_resize_u8, load_panorama_batch
    z = np.load(os.path.join(golden_dir, "area_resize.npz"))
    src = torch.from_numpy(z["src"]).to(DEV)
    for f in (2, 4, 8):
        got = area_resize_u8(torch.stack([src, src.flip(0)]), f).cpu().numpy()
        assert np.array_equal(got[0], z[f"area_{f}"]), f
    x = load_panorama_batch(src, 2).cpu().numpy()
    assert np.array_equal(x[0], (z["area_2"].astype(np.float32) / 255).transpose(2, 0, 1))
    with pytest.raises(Exception):
        area_resize_u8(src, 5)                           # 64 % 5 != 0


@pytest.mark.parametrize("n,keep", [(2 * 128 * 256, 0.8), (8 * 512 * 1024, 0.5), (1027, 1.0), (4096, 0.0)])
def test_masked_median_matches_torch_median(n, keep):
    """test.py:161-162 uses torch.median (the LOWER median) of the masked values; the radix-select kernel must return
    exactly that element, and the scale factor the same float32 quotient."""
    from omnifusion_b200 import metrics
    g = torch.Generator().manual_seed(n)
    pred = (0.05 + 9.0 * torch.rand(n, generator=g)).view(1, 1, 1, n)
    gt = (0.1 + 7.9 * torch.rand(n, generator=g)).view(1, 1, 1, n)
    gt[0, 0, 0, : n // 7] = gt[0, 0, 0, 0]              # a long run of equal values around which ranks are ambiguous
    mask = (torch.rand(n, generator=g) < keep).view(1, 1, 1, n)
    out = metrics.median_scale_device(pred.to(DEV), gt.to(DEV), mask.to(DEV)).cpu()
    if mask.any():
        mg, mp = gt[mask].median(), pred[mask].median()
        assert out[1].item() == mg.item() and out[2].item() == mp.item()
        assert out[0].item() == (mg / mp).item()
    else:
        assert torch.isnan(out[1]) and torch.isnan(out[2])


def test_depth_meters_follow_the_reference_average_meters():
    """test.py:121-177: seven AverageMeters updated with (per-batch value, N = mask.sum()); rms_sq_log's per-batch
    value is a mean over ITS OWN valid pixels (pred > 1e-7), so it is not a plain ratio of global sums."""
    from omnifusion_b200 import metrics
    from oracle import model as om
    meters = metrics.DepthMeters(DEV)
    want = {k: 0.0 for k in metrics.METRIC_NAMES}
    count = 0
    for b in range(3):
        pred = 0.1 + 7.9 * urand(2, 1, 64, 128, seed=10 + b)
        pred[0, 0, b, :50] = 0.0                         # pred <= 1e-7: excluded from rms_sq_log only
        gt = 0.1 + 7.9 * urand(2, 1, 64, 128, seed=20 + b)
        mask = (gt <= 8) & (gt > 0.1) & (urand(2, 1, 64, 128, seed=30 + b) > 0.1 * (b + 1))
        ref = om.eval_metrics(pred, gt, mask, median_scale=True)
        for k in want:
            want[k] += ref[k] * ref["n"]
        count += ref["n"]
        meters.update(pred.to(DEV), gt.to(DEV), mask.to(DEV), use_median_scale=True)
    got = meters.all_reduce().result()
    assert got["n"] == count
    for k in want:
        assert abs(got[k] - want[k] / count) <= 2e-6 * max(1.0, abs(want[k] / count)), (k, got[k], want[k] / count)
    with pytest.raises(ValueError):
        metrics.depth_metrics_partial(torch.zeros(2, 1, 4, 4, device=DEV), torch.zeros(2, 4, 4, device=DEV),
                                      torch.ones(2, 1, 4, 4, device=DEV))


@pytest.mark.parametrize("nrows,erp,P,B", [(4, (128, 256), 32, 5), (6, (64, 128), 16, 3), (4, (512, 1024), 128, 2)])
def test_blend_conf_pair_map_matches_separate_maps_and_oracle(nrows, erp, P, B):
    """ofb_blend_conf_pairs_f32 (interleaved (pred*conf, conf) map, the engine's hot path) against the two-map kernel
    (bit-identical: same table walk, same expression) and the oracle; nrows=6 at a tiny ERP makes CSR rows long
    enough to overflow the shared-memory stage (global-memory fallback for the entries beyond the capacity)."""
    o = ops()
    n = tables.NUM_PATCHES[nrows]
    pred = urand(B * n, P, P, seed=1) * 3
    conf = urand(B * n, P, P, seed=2)
    tab = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in tables.blend_table(FOV, nrows, (P, P), erp).items()}
    sep = o.blend_conf((pred * conf).to(DEV), conf.to(DEV), B, n, tab, *erp)
    pairs = torch.stack([pred * conf, conf], -1).contiguous().to(DEV)
    got = torch.empty((B, 1, *erp), device=DEV)
    _lib.check(_lib.lib().ofb_blend_conf_pairs_f32(_lib.ptr(pairs), B, n, P, P, _lib.ptr(tab["rowptr"]), _lib.ptr(tab["idx"]),
                                                   _lib.ptr(tab["w"]), erp[0], erp[1], _lib.ptr(got),
                                                   _lib.stream_of(torch.device(DEV))))
    torch.cuda.synchronize()
    assert torch.equal(got, sep)
    unf = lambda t: t.reshape(B, n, 1, P, P).permute(0, 2, 3, 4, 1)
    W = oe.pers2equi(unf(conf), FOV, nrows, (P, P), erp)
    D = oe.pers2equi(unf(pred * conf), FOV, nrows, (P, P), erp)
    report("blend_conf pairs", got.cpu(), D / (W + 1e-8 * (W <= 1e-8).float()), atol=0, rtol=1e-5)
    rows = (tab["rowptr"][1:] - tab["rowptr"][:-1]).reshape(-1)
    print(f"[parity] blend table nrows={nrows} erp={erp}: max entries per 256-pixel tile = "
          f"{int(rows.cpu().reshape(-1, 256).sum(1).max()) if rows.numel() % 256 == 0 else -1}")


@pytest.mark.parametrize("nrows,erp,P", [(4, (64, 128), 16), (6, (128, 256), 32), (3, (32, 64), 16)])
def test_resampler_backward_matches_autograd_of_the_oracle(nrows, erp, P):
    """Training direction: gradients of equi2pers w.r.t. the ERP image and of pers2equi w.r.t. the patches
    against torch autograd through the CPU restatement (F.grid_sample / gather + weighted sum, as the reference).
    Scatter-adds are atomic (order not deterministic): tolerance 1e-5 relative to the gradient's magnitude."""
    from omnifusion_b200.equi_pers.equi2pers_v3 import equi2pers
    from omnifusion_b200.equi_pers.pers2equi_v3 import pers2equi
    n = tables.NUM_PATCHES[nrows]
    img = urand(2, 3, *erp, seed=31).requires_grad_(True)
    go = rand(2, 3, P, P, n, seed=32)
    ref_p = oe.equi2pers(img, FOV, nrows, (P, P))[0]
    ref_p.backward(go)
    img_d = img.detach().to(DEV).requires_grad_(True)
    got_p = equi2pers(img_d, FOV, nrows, (P, P))[0]
    got_p.backward(go.to(DEV))
    err = (img_d.grad.cpu() - img.grad).abs().max().item() / img.grad.abs().max().item()
    print(f"[parity] equi2pers backward nrows={nrows}: max err / max |grad| = {err:.2e}")
    assert err <= 1e-5
    pers = urand(2, 2, P, P, n, seed=33).requires_grad_(True)
    ge = rand(2, 2, *erp, seed=34)
    oe.pers2equi(pers, FOV, nrows, (P, P), erp).backward(ge)
    pers_d = pers.detach().to(DEV).requires_grad_(True)
    pers2equi(pers_d, FOV, nrows, (P, P), erp, "x").backward(ge.to(DEV))
    err = (pers_d.grad.cpu() - pers.grad).abs().max().item() / pers.grad.abs().max().item()
    print(f"[parity] pers2equi backward nrows={nrows}: max err / max |grad| = {err:.2e}")
    assert err <= 1e-5
