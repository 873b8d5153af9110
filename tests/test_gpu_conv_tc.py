"""tcgen05 implicit-GEMM conv engine vs torch-CPU fp32 (and vs the CUDA-core engine).

F16X3 mode (split-half planes, three kind::f16 MMAs) must be fp32-accurate; TF32 mode (one
kind::tf32 MMA on float32 tensors) is only TF32-accurate and is tested at 3e-3."""
import pytest
import torch
import torch.nn.functional as F

from omnifusion_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ops():
    import ofb_ops
    return ofb_ops


def rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


TC_CASES = [
    # n, hw, c0, c1, cout, k, residual, act [, stride]
    (2, 32, 64, 0, 64, 3, True, 1),       # layer1 block conv
    (3, 16, 128, 0, 128, 3, True, 1),     # layer2
    (5, 8, 256, 0, 256, 3, False, 1),     # layer3 (2 images per tile, ragged last group)
    (19, 4, 512, 0, 512, 3, True, 1),     # layer4 (8 images per tile, ragged)
    (3, 8, 256, 256, 128, 3, False, 1),   # de_conv0_1 (concat)
    (2, 64, 64, 64, 32, 3, False, 1),     # de_conv3_1 (concat, cout 32)
    (1, 128, 32, 0, 32, 3, False, 1),     # de_conv4_0 (cin 32: 64-byte rows in F16X3)
    (2, 64, 64, 0, 64, 3, False, 1),      # de_conv3_0
    (200, 1, 512, 0, 2048, 1, False, 2),  # fc1 + GELU as 1x1
    (144, 1, 2048, 0, 512, 1, True, 0),   # fc2 + residual
    (4, 4, 512, 0, 32, 1, False, 0),      # down1
    (3, 32, 64, 0, 128, 3, False, 1, 2),  # layer2.0.conv1 (3x3 stride 2 via TMA traversal strides)
    (3, 32, 64, 0, 128, 1, False, 0, 2),  # layer2.0.downsample (1x1 stride 2)
    (5, 8, 256, 0, 512, 3, False, 1, 2),  # layer4.0.conv1
    (5, 8, 256, 0, 512, 1, False, 0, 2),  # layer4.0.downsample
    (3, 16, 128, 128, 64, 3, False, 1),   # de_conv1_1 (kh-reuse tiling, 16x8 pixel tiles)
    (2, 32, 64, 64, 64, 3, True, 1),      # de_conv2_1-like with residual (kh-reuse, 32x4 tiles)
]

# Shapes with >= 74 (= SMs / 2) tiles of 128 output channels: conv_tc keeps BN = 128 and launches the cta_group::2
# CTA-pair instantiation (conv_tc_kernel<128, F16X3, 128, .., CTA2>), the dominant kernel class of the benchmark
# step.  (The small TC_CASES above all shrink to BN = 32 / 64 and never reach it.)
CTA2_CASES = [
    (150, 8, 256, 0, 256, 3, True, 1),     # layer3 block conv at B=8-like size: 75 M tiles (odd: ragged last pair) x 2 N tiles
    (600, 4, 512, 0, 512, 3, True, 1),     # layer4: 8 images per tile, 75 M tiles x 4 N tiles, K = 4608
    (40, 16, 128, 0, 128, 3, False, 1),    # layer2: half-image tiles, 80 M tiles x 1 N tile
    (150, 8, 256, 256, 256, 3, False, 1),  # de_conv0_0-like with a concat source (two tensor maps), K = 4608
    (150, 16, 128, 0, 256, 3, False, 1, 2),  # layer3.0.conv1: 3x3 stride 2 (TMA traversal strides) on CTA pairs
    (150, 16, 128, 0, 256, 1, False, 0, 2),  # layer3.0.downsample: 1x1 stride 2, K = 128 (two K-steps)
    (149, 8, 256, 0, 128, 3, False, 1),    # odd image count: the last tile holds one image, 75 M tiles x 1 N tile
]


def reference(case, seed=0):
    n, hw, c0, c1, cout, k, use_res, act = case[:8]
    stride = case[8] if len(case) > 8 else 1
    x0 = rand(n, c0, hw, hw, seed=seed + 1)
    x1 = rand(n, c1, hw, hw, seed=seed + 2) if c1 else None
    w = rand(cout, c0 + c1, k, k, seed=seed + 3, scale=(1.0 / ((c0 + c1) * k * k)) ** 0.5)
    scale = 0.5 + torch.rand(cout, generator=torch.Generator().manual_seed(seed + 4))
    shift = rand(cout, seed=seed + 5, scale=0.1)
    res = rand(n, cout, hw // stride, hw // stride, seed=seed + 6) if use_res else None
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    y = F.conv2d(x, w, None, stride, k // 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if res is not None:
        y = y + res
    y = F.relu(y) if act == 1 else (F.gelu(y) if act == 2 else y)
    return x0, x1, w, scale, shift, res, y


def run(case, engine, fmt):
    o = ops()
    n, hw, c0, c1, cout, k, use_res, act = case[:8]
    stride = case[8] if len(case) > 8 else 1
    x0, x1, w, scale, shift, res, y = reference(case)
    got = o.conv_fmt(o.nhwc(x0).to(DEV), o.ohwi(w).to(DEV), k, stride, k // 2,
                     in1=o.nhwc(x1).to(DEV) if c1 else None, scale=scale.to(DEV), shift=shift.to(DEV),
                     residual=o.nhwc(res).to(DEV) if use_res else None, act=act, engine=engine, in_fmt=fmt, out_fmt=fmt)
    torch.cuda.synchronize()
    return o.nchw(got.cpu()), y


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_f16x3_matches_fp32(case):
    got, ref = run(case, _lib.ENGINE_TC, _lib.FMT_SPLIT16)
    err = (got - ref).abs()
    print(f"[parity] conv_tc f16x3 {case}: max_abs_err={err.max().item():.3e} ref_absmax={ref.abs().max().item():.3e}")
    # tensor-core fp32 accumulation truncates (round-toward-zero) on every K-block add, so the error
    # grows ~linearly with K (observed 1e-4 abs at K=4608) instead of ~sqrt(K) for FFMA
    assert (err <= 1.5e-4 + 5e-5 * ref.abs()).all()


@pytest.mark.parametrize("case", CTA2_CASES)
def test_conv_tc_cta_pair_kernel_matches_fp32(case):
    """The cta_group::2 instantiation directly against torch-CPU fp32, and bit-identical to the single-CTA
    instantiation of the same layer (option cta2=0 keeps the same `group64` accumulation grouping only when both
    run with it, so the comparison is against fp32 for both and between them at rounding level)."""
    o = ops()
    before = _lib.lib().ofb_launch_count(1)
    got, ref = run(case, _lib.ENGINE_TC, _lib.FMT_SPLIT16)
    err = (got - ref).abs()
    print(f"[parity] conv_tc cta_group::2 {case}: max_abs_err={err.max().item():.3e} ref_absmax={ref.abs().max().item():.3e}")
    assert (err <= 1.5e-4 + 5e-5 * ref.abs()).all()
    assert o.last_conv_variant() == "cta2", "this shape must select the CTA-pair kernel"


@pytest.mark.parametrize("case", TC_CASES[:7])
def test_conv_tc_tf32_matches_fp32_loosely(case):
    got, ref = run(case, _lib.ENGINE_TC, _lib.FMT_F32)
    err = (got - ref).abs()
    print(f"[parity] conv_tc tf32 {case}: max_abs_err={err.max().item():.3e} ref_absmax={ref.abs().max().item():.3e}")
    assert (err <= 1e-2 + 3e-3 * ref.abs()).all()


@pytest.mark.parametrize("case", [TC_CASES[0], TC_CASES[4], TC_CASES[9]])
def test_conv_simt_split_format_matches_fp32(case):
    got, ref = run(case, _lib.ENGINE_SIMT, _lib.FMT_SPLIT16)
    err = (got - ref).abs()
    print(f"[parity] conv_simt split {case}: max_abs_err={err.max().item():.3e}")
    assert (err <= 3e-5 + 3e-5 * ref.abs()).all()


def test_split_merge_roundtrip_precision():
    o = ops()
    x = rand(1 << 16, seed=1, scale=10.0).to(DEV)
    back = o.merge16(o.split16(x), x.shape)
    err = (back - x).abs()
    print(f"[parity] split16 roundtrip max abs err {err.max().item():.3e}")
    # ~22 mantissa bits, and an absolute floor of half an fp16 subnormal step (2^-25) on the lo plane
    assert (err <= 3.1e-8 + 2.5e-7 * x.abs()).all()


def test_stem_on_tensor_cores_matches_fp32():
    """7x7 s2 stem as an implicit GEMM over overlapping 64-byte TMA windows (ofb_stem_tc_f16)."""
    import ctypes as C
    o = ops()
    n = 3
    x = torch.rand(n, 3, 128, 128, generator=torch.Generator().manual_seed(1))
    w = rand(64, 3, 7, 7, seed=2, scale=0.1)
    scale = 0.5 + torch.rand(64, generator=torch.Generator().manual_seed(3))
    shift = rand(64, seed=4, scale=0.1)
    ref = F.relu(F.conv2d(x, w, None, 2, 3) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    # patches: (n,128,128+8,4) with zero row pads and a zero 4th channel, split-half planes
    xp = torch.zeros(n, 128, 136, 4)
    xp[:, :, 4:132, :3] = x.permute(0, 2, 3, 1)
    planes = o.split16(xp.to(DEV))
    # weights: (64,7,8,4) with kw' = kw + 1
    wq = torch.zeros(64, 7, 8, 4)
    wq[:, :, 1:, :3] = w.permute(0, 2, 3, 1)
    mul = o.weight_scale(wq)
    wplanes = o.split16(wq.to(DEV), mul)
    out = torch.empty(2 * n * 64 * 64 * 64, dtype=torch.float16, device=DEV)
    d_scale, d_shift = scale.to(DEV), shift.to(DEV)          # keep the device copies alive across the call
    _lib.use_device(torch.device(DEV))
    _lib.check(_lib.lib().ofb_stem_tc_f16(_lib.ptr(planes), n, 128, 128, _lib.ptr(wplanes), 1.0 / mul,
                                           _lib.ptr(d_scale), _lib.ptr(d_shift), _lib.ptr(out),
                                           _lib.stream_of(torch.device(DEV))))
    torch.cuda.synchronize()
    got = o.nchw(o.merge16(out, (n, 64, 64, 64)).cpu())
    err = (got - ref).abs()
    print(f"[parity] stem_tc: max_abs_err={err.max().item():.3e} ref_absmax={ref.abs().max().item():.3e}")
    assert (err <= 3e-5 + 3e-5 * ref.abs()).all()


@pytest.mark.parametrize("n,hw", [(1, 64), (3, 64), (150, 64)])
def test_conv_tc_fused_upsample_matches_interpolate_then_conv(n, hw):
    """de_conv4_0 with the decoder's last F.interpolate(x2, bilinear, align_corners=False) folded into the
    conv's operand producer (ofb_conv_desc.ups2x): against torch-CPU fp32, and against libofb's own
    unfused upsample2x kernel + conv (same expression tree for the interpolation).  The fused kernel
    walks contiguous ranges of 128-pixel output rows per CTA: n=1 gives CTAs one row each (every row is a
    segment with both halo rows re-produced), n=3 makes ranges cross image boundaries, n=150 exceeds the
    ring of upsampled rows many times over."""
    o = ops()
    x = rand(n, 32, hw, hw, seed=11)
    w = rand(32, 32, 3, 3, seed=12, scale=(1.0 / (32 * 9)) ** 0.5)
    scale = 0.5 + torch.rand(32, generator=torch.Generator().manual_seed(13))
    shift = rand(32, seed=14, scale=0.1)
    up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    ref = F.relu(F.conv2d(up, w, None, 1, 1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    xd, wd, sd, td = o.nhwc(x).to(DEV), o.ohwi(w).to(DEV), scale.to(DEV), shift.to(DEV)
    fused = o.conv_fmt(xd, wd, 3, 1, 1, scale=sd, shift=td, act=1, engine=_lib.ENGINE_TC,
                       in_fmt=_lib.FMT_SPLIT16, out_fmt=_lib.FMT_SPLIT16, ups2x=1)
    # unfused: upsample kernel on split-half planes, then the same conv kernel family
    up_planes = torch.empty(2 * n * 4 * hw * hw * 32, dtype=torch.float16, device=DEV)
    xs = o.split16(xd)
    _lib.check(_lib.lib().ofb_upsample2x_f32(_lib.ptr(xs), None, n, hw, hw, 32, _lib.ptr(up_planes), 1,
                                              _lib.stream_of(torch.device(DEV))))
    up_f32 = o.merge16(up_planes, (n, 2 * hw, 2 * hw, 32))
    unfused = o.conv_fmt(up_f32, wd, 3, 1, 1, scale=sd, shift=td, act=1, engine=_lib.ENGINE_TC,
                         in_fmt=_lib.FMT_SPLIT16, out_fmt=_lib.FMT_SPLIT16)
    torch.cuda.synchronize()
    got = o.nchw(fused.cpu())
    err = (got - ref).abs()
    d2 = (fused - unfused).abs().max().item()
    print(f"[parity] conv_tc fused upsample n={n} {hw}->{2 * hw}: max_abs_err={err.max().item():.3e} "
          f"ref_absmax={ref.abs().max().item():.3e} vs unfused libofb path max diff {d2:.3e}")
    # the tap-stacked variant of the same kernel (option nstack = 2; used by the heads, kept for experiments here)
    _lib.check(_lib.lib().ofb_debug_nstack(2))
    stacked = o.conv_fmt(xd, wd, 3, 1, 1, scale=sd, shift=td, act=1, engine=_lib.ENGINE_TC,
                         in_fmt=_lib.FMT_SPLIT16, out_fmt=_lib.FMT_SPLIT16, ups2x=1)
    _lib.check(_lib.lib().ofb_debug_nstack(1))
    torch.cuda.synchronize()
    assert ((o.nchw(stacked.cpu()) - ref).abs() <= 1.5e-4 + 5e-5 * ref.abs()).all()
    assert (err <= 1.5e-4 + 5e-5 * ref.abs()).all()
    # same products, different summation order (the rolling-row kernel accumulates an output row input row by input
    # row, the tile-box kernel tap by tap): agreement to fp32 rounding of values up to ~4
    assert d2 <= 5e-6


def test_conv_fused_upsample_rejected_by_cuda_core_engine():
    o = ops()
    x = o.nhwc(rand(1, 32, 16, 16, seed=1)).to(DEV)
    w = o.ohwi(rand(32, 32, 3, 3, seed=2)).to(DEV)
    with pytest.raises(_lib.OfbError):
        o.conv_fmt(x, w, 3, 1, 1, engine=_lib.ENGINE_SIMT, in_fmt=_lib.FMT_SPLIT16, out_fmt=_lib.FMT_SPLIT16, ups2x=1)


@pytest.mark.parametrize("n,h,confidence", [(1, 128, True), (3, 128, False), (150, 16, True)])
def test_heads_on_tensor_cores_match_torch_cpu(n, h, confidence):
    """pred / weight_pred heads (spherical_model_iterative.py:371-374) as one 16-channel conv on the rolling-row
    tcgen05 kernel (ofb_heads_tc_f16): against torch-CPU fp32.  n=1 gives every CTA one row (both halo rows
    re-loaded per row), n=3 makes CTA ranges cross image boundaries, n=150 wraps the row ring many times."""
    o = ops()
    x = F.relu(rand(n, 32, h, 128, seed=21))
    wp, wc = rand(1, 32, 3, 3, seed=22, scale=0.1), rand(1, 32, 3, 3, seed=23, scale=0.1)
    bp, bc = 0.7, -0.3
    pred = F.relu(F.conv2d(x, wp, torch.tensor([bp]), 1, 1))
    conf = torch.sigmoid(F.conv2d(x, wc, torch.tensor([bc]), 1, 1))
    w16 = torch.zeros(16, 3, 3, 32)
    w16[0], w16[1] = o.ohwi(wp)[0], o.ohwi(wc)[0]
    w16 = w16.to(DEV)
    mul = o.weight_scale(w16)
    ws, xs = o.split16(w16, mul), o.split16(o.nhwc(x).to(DEV))
    gp = torch.full((n, h, 128), -1.0, device=DEV)
    gc = torch.full((n, h, 128), -1.0, device=DEV)
    _lib.check(_lib.lib().ofb_heads_tc_f16(_lib.ptr(xs), n, h, 128, _lib.ptr(ws), 1.0 / mul, bp, bc, int(confidence),
                                            _lib.ptr(gp), _lib.ptr(gc), _lib.stream_of(torch.device(DEV))))
    torch.cuda.synchronize()
    want = (pred * conf)[:, 0] if confidence else pred[:, 0]
    err = (gp.cpu() - want).abs()
    print(f"[parity] heads_tc n={n} h={h} conf={confidence}: max_abs_err={err.max().item():.3e} ref_absmax={want.abs().max().item():.3e}")
    assert (err <= 2e-5 + 2e-5 * want.abs()).all()
    if confidence:
        assert ((gc.cpu() - conf[:, 0]).abs() <= 2e-5).all()
        # the engine's layout: (pred*conf, conf) interleaved per pixel - must carry exactly the same numbers
        pairs = torch.full((n, h, 128, 2), -1.0, device=DEV)
        _lib.check(_lib.lib().ofb_heads_tc_pairs_f16(_lib.ptr(xs), n, h, 128, _lib.ptr(ws), 1.0 / mul, bp, bc,
                                                      _lib.ptr(pairs), _lib.stream_of(torch.device(DEV))))
        torch.cuda.synchronize()
        assert torch.equal(pairs[..., 0], gp) and torch.equal(pairs[..., 1], gc)
    else:
        assert (gc == -1.0).all()          # conf_out untouched


@pytest.mark.parametrize("rows,cin,ksplit", [(144, 2048, 4), (18, 512, 4), (576, 512, 2)])
def test_splitk_linear_and_finish_layernorm(rows, cin, ksplit):
    """attn.proj / mlp.fc2 of a Transformer_Block (model/blocks.py:84-88) as a split-K linear on the tcgen05 engine
    (ofb_conv_desc.ksplit) + ofb_splitk_finish_ln_f32 (bias + residual + the following LayerNorm), against
    torch-CPU fp32."""
    o = ops()
    x = rand(rows, cin, seed=31)
    w = rand(512, cin, seed=32, scale=(1.0 / cin) ** 0.5)
    bias, res = rand(512, seed=33, scale=0.1), rand(rows, 512, seed=34)
    gamma, beta = 0.5 + torch.rand(512, generator=torch.Generator().manual_seed(35)), rand(512, seed=36, scale=0.1)
    want_x = res + F.linear(x, w, bias)
    want_ln = F.layer_norm(want_x, (512,), gamma, beta, 1e-5)
    xd = x.view(rows, 1, 1, cin).to(DEV)
    wd = w.view(512, 1, 1, cin).to(DEV)
    part, unscale = o.conv_fmt(xd, wd, 1, 1, 0, engine=_lib.ENGINE_TC, in_fmt=_lib.FMT_SPLIT16,
                               out_fmt=_lib.FMT_SPLIT16, ksplit=ksplit)
    resd = o.split16(res.to(DEV))
    bias_d, gamma_d, beta_d = bias.to(DEV), gamma.to(DEV), beta.to(DEV)      # kept alive across the launches
    x_out = torch.empty(2 * rows * 512, dtype=torch.float16, device=DEV)
    for ln_fmt in (1, 0):
        ln_out = torch.empty(2 * rows * 512, dtype=torch.float16, device=DEV) if ln_fmt else torch.empty(rows, 512, device=DEV)
        _lib.check(_lib.lib().ofb_splitk_finish_ln_f32(_lib.ptr(part), ksplit, unscale, _lib.ptr(bias_d), _lib.ptr(resd),
                                                        rows, 512, _lib.ptr(x_out), _lib.ptr(gamma_d),
                                                        _lib.ptr(beta_d), 1e-5, _lib.ptr(ln_out), ln_fmt,
                                                        _lib.stream_of(torch.device(DEV))))
        torch.cuda.synchronize()
        got_x = o.merge16(x_out, (rows, 512)).cpu()
        got_ln = (o.merge16(ln_out, (rows, 512)) if ln_fmt else ln_out).cpu()
        ex, el = (got_x - want_x).abs().max().item(), (got_ln - want_ln).abs().max().item()
        print(f"[parity] split-K linear rows={rows} K={cin} S={ksplit} ln_fmt={ln_fmt}: x max_abs_err={ex:.3e} ln max_abs_err={el:.3e}")
        assert ex <= 2e-5 and el <= 5e-5


@pytest.mark.parametrize("n,cin,cout,hw", [(144, 512, 512, 4), (18, 512, 512, 4), (36, 256, 256, 8)])
def test_splitk_conv3x3_and_finish(n, cin, cout, hw):
    """layer4's 512 -> 512 BasicBlock convs (torchvision resnet34 via spherical_model_iterative.py:328) as a 2-way
    split-K 3x3 conv on the tcgen05 engine (CTA-pair kernel, each slice = half of the channel chunks of every tap) +
    ofb_splitk_finish_conv_f16 (BN scale / shift, residual, ReLU, split-half store), against torch-CPU fp32."""
    o = ops()
    x = rand(n, hw, hw, cin, seed=71)
    w = rand(cout, 3, 3, cin, seed=72, scale=(1.0 / (9 * cin)) ** 0.5)
    scale = 0.5 + torch.rand(cout, generator=torch.Generator().manual_seed(73))
    shift, res = rand(cout, seed=74, scale=0.1), rand(n, hw, hw, cout, seed=75)
    want = F.relu(F.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), padding=1) * scale.view(1, -1, 1, 1)
                  + shift.view(1, -1, 1, 1) + res.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    part, unscale = o.conv_fmt(x.to(DEV), w.to(DEV), 3, 1, 1, engine=_lib.ENGINE_TC, in_fmt=_lib.FMT_SPLIT16,
                               out_fmt=_lib.FMT_SPLIT16, ksplit=2)
    variant = o.last_conv_variant()
    resd, sc_d, sh_d = o.split16(res.to(DEV)), scale.to(DEV), shift.to(DEV)
    out = torch.full((2 * n * hw * hw * cout,), float("nan"), dtype=torch.float16, device=DEV)
    _lib.check(_lib.lib().ofb_splitk_finish_conv_f16(_lib.ptr(part), 2, n * hw * hw, cout, _lib.ptr(sc_d), _lib.ptr(sh_d),
                                                      unscale, _lib.ptr(resd), 1, _lib.ptr(out),
                                                      _lib.stream_of(torch.device(DEV))))
    torch.cuda.synchronize()
    got = o.merge16(out, (n, hw, hw, cout)).cpu()
    err = (got - want).abs()
    print(f"[parity] split-K conv3x3 n={n} {cin}->{cout} @{hw}x{hw} variant={variant}: max_abs_err={err.max().item():.3e} "
          f"ref_absmax={want.abs().max().item():.3e}")
    assert torch.isfinite(got).all()
    assert (err <= 1.5e-4 + 5e-5 * want.abs()).all()
    # the two slices really are halves of K: each partial sum alone differs from the total
    assert (part[0] - part[1]).abs().max().item() > 0


@pytest.mark.parametrize("B,n_tok", [(8, 18), (3, 18), (16, 46), (5, 26), (13, 10), (1, 64)])
def test_attention_on_tensor_cores_matches_torch_cpu(B, n_tok):
    """Attention core of model/blocks.py:50-62 on tcgen05 (ofb_attention_tc_f16): one CTA per head and tile of
    128 // N whole panoramas, block-diagonal softmax.  Against torch-CPU fp32; B chosen so that the last tile is
    ragged (8 = 7 + 1 panoramas at N = 18) and so that N = 64 fills half a tile exactly."""
    o = ops()
    heads, d = 4, 128
    qkv = rand(B * n_tok, 3 * heads * d, seed=41 + n_tok)
    q, k, v = qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:]
    sh = lambda t: t.reshape(B, n_tok, heads, d).permute(0, 2, 1, 3)
    att = ((sh(q) @ sh(k).transpose(-2, -1)) * d ** -0.5).softmax(-1)
    ref = (att @ sh(v)).transpose(1, 2).reshape(B * n_tok, heads * d)
    planes = o.split16(qkv.to(DEV))
    out = torch.full((2 * B * n_tok * heads * d,), float("nan"), dtype=torch.float16, device=DEV)
    _lib.check(_lib.lib().ofb_attention_tc_f16(_lib.ptr(planes), B, n_tok, heads, d, _lib.ptr(out),
                                                _lib.stream_of(torch.device(DEV))))
    torch.cuda.synchronize()
    got = o.merge16(out, (B * n_tok, heads * d)).cpu()
    err = (got - ref).abs()
    print(f"[parity] attention_tc B={B} N={n_tok}: max_abs_err={err.max().item():.3e} ref_absmax={ref.abs().max().item():.3e}")
    assert torch.isfinite(got).all()
    assert (err <= 3e-6 + 1e-5 * ref.abs()).all()
