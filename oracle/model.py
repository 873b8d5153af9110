"""Oracle restatement of the OmniFusion networks (TEST INFRASTRUCTURE ONLY).

A functional torch-CPU fp32 forward driven directly by a reference-layout
``state_dict``.  Follows

* iterative model  - /root/reference/model/spherical_model_iterative.py:308-456
* single-stage     - /root/reference/model/spherical_model.py:238-314
* transformer      - /root/reference/model/blocks.py:14-89 and
                     spherical_model_iterative.py:232-250
* encoder topology - torchvision ResNet-34 (BasicBlock x 3,4,6,3) as rewritten
                     by convert_conv / convert_bn (spherical_model_iterative.py:185-230)

Restatement choice: the reference runs every conv as Conv3d((k,k,1)) over
(B,C,H,W,N); here the patch axis is folded into the batch, image index b*N+n,
and the same weights (trailing unit dim dropped) are applied with conv2d over
(B*N,C,H,W).  That is the same sum of products in a different loop order; the
measured difference to the real reference is recorded by make_golden.py in
tests/golden/meta.json (max rel ~1e-6).
"""
import torch
import torch.nn.functional as F

from .equi_pers import equi2pers, pers2equi


def _pair(t):
    return t if isinstance(t, tuple) else (t, t)


def strip_module_prefix(sd):
    """test.py:107-110 saves DataParallel checkpoints whose keys start with 'module.'.
    Also drops the trailing unit dim of the Conv3d((k,k,1)) weights once, so conv2d
    gets contiguous 4-D filters."""
    if all(k.startswith("module.") for k in sd):
        sd = {k[len("module."):]: v for k, v in sd.items()}
    return {k: (v[..., 0].contiguous() if v.dim() == 5 else v) for k, v in sd.items()}


def _w2d(sd, name):
    return sd[name]


def _bn(sd, prefix, x, eps=1e-5):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, eps)


def _conv(sd, name, x, stride=1, pad=0):
    return F.conv2d(x, _w2d(sd, name + ".weight"), sd.get(name + ".bias"), stride, pad)


def _basic_block(sd, p, x, stride):
    out = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x, stride, 1)))
    out = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out, 1, 1))
    if (p + ".downsample.0.weight") in sd:
        x = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x, stride, 0))
    return F.relu(out + x)


def _res_layer(sd, name, x, nblocks, stride):
    for i in range(nblocks):
        x = _basic_block(sd, f"{name}.{i}", x, stride if i == 0 else 1)
    return x


def _cbr(sd, name, x):
    """ConvBnReLU_v2 (spherical_model_iterative.py:29-37)."""
    return F.relu(_bn(sd, name + ".bn", _conv(sd, name + ".conv", x, 1, 1)))


def _mlp_points(sd, name, x):
    """1x1 conv, BN, ReLU, 1x1 conv, BN, ReLU (spherical_model_iterative.py:290-305)."""
    x = F.relu(_bn(sd, name + ".1", F.conv2d(x, sd[name + ".0.weight"])))
    return F.relu(_bn(sd, name + ".4", F.conv2d(x, sd[name + ".3.weight"])))


def _attention(sd, p, x, heads=4):
    """blocks.py:50-66."""
    B, N, C = x.shape
    q = F.linear(x, sd[p + ".q.weight"]).reshape(B, N, heads, C // heads).permute(0, 2, 1, 3)
    kv = F.linear(x, sd[p + ".kv.weight"]).reshape(B, -1, 2, heads, C // heads).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    attn = (q @ k.transpose(-2, -1)) * ((C // heads) ** -0.5)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def transformer(sd, tokens, depth=6, heads=4, trace=None):
    """Transformer_cascade.forward (spherical_model_iterative.py:243-250) over
    Transformer_Block (blocks.py:84-88).  tokens (B,N,512)."""
    C = tokens.shape[-1]
    h = tokens + sd["transformer.pos_emb"]
    for i in range(depth):
        p = f"transformer.layer.{i}"
        y = F.layer_norm(h, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)
        h = h + _attention(sd, p + ".attn", y, heads)
        y = F.layer_norm(h, (C,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5)
        y = F.gelu(F.linear(y, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"]))
        h = h + F.linear(y, sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
        if trace is not None:
            trace[f"block{i}"] = h
    return F.layer_norm(h, (C,), sd["transformer.encoder_norm.weight"],
                        sd["transformer.encoder_norm.bias"], 1e-6)


def _fold(x):
    """(B,C,H,W,N) -> (B*N,C,H,W)."""
    B, C, H, W, N = x.shape
    return x.permute(0, 4, 1, 2, 3).reshape(B * N, C, H, W)


def _unfold(x, B):
    """(B*N,C,H,W) -> (B,C,H,W,N)."""
    BN, C, H, W = x.shape
    return x.reshape(B, BN // B, C, H, W).permute(0, 2, 3, 4, 1)


def patch_network(sd, patches, point_feat, bs, down_name="down1", trace=None, use_points=True, use_transformer=True):
    """Encoder + token fusion + decoder + both heads on folded patches.

    patches (B*N,3,P,P); point_feat (B*N,64,P/4,P/4).
    Returns (pred_raw, weight_raw, de_conv4_0): the two 3x3 head convs *before*
    relu / sigmoid.  spherical_model_iterative.py:322-371.
    """
    n_patch = patches.shape[0] // bs
    tr = trace if trace is not None else {}
    conv1 = F.relu(_bn(sd, "bn1", _conv(sd, "conv1", patches, 2, 3)))
    pool = F.max_pool2d(conv1, kernel_size=3, stride=2, padding=1)
    layer1 = _res_layer(sd, "layer1", pool, 3, 1)
    tr["layer1_pre"] = layer1
    if use_points:                                             # network_360d.py:325 comments this add out
        layer1 = layer1 + point_feat
    layer2 = _res_layer(sd, "layer2", layer1, 4, 2)
    layer3 = _res_layer(sd, "layer3", layer2, 6, 2)
    layer4 = _res_layer(sd, "layer4", layer3, 3, 2)
    tr.update(conv1=conv1, pool=pool, layer1=layer1, layer2=layer2, layer3=layer3, layer4=layer4)

    # token = flattened (c, i, j) of the 32x4x4 map (spherical_model_iterative.py:330-331)
    if use_transformer:                                        # network_360d.py:330-335 comments the token path out
        down = _conv(sd, down_name, layer4)                       # (B*N,32,4,4); (B*N,8,8,8) for 256x256 patches
        tokens = down.reshape(bs, n_patch, -1)                     # (B,N,512)
        tr["tokens"] = tokens
        enc = transformer(sd, tokens, trace=tr)                    # (B,N,512)
        tr["encoded"] = enc
        layer4 = layer4 + enc.reshape(bs * n_patch, -1, 1, 1)      # broadcast over the SxS map (:334-335)

    def up(x, ref):
        return F.interpolate(x, size=ref.shape[-2:], mode="bilinear", align_corners=False)

    d00 = _cbr(sd, "de_conv0_0", up(layer4, layer3))
    d01 = _cbr(sd, "de_conv0_1", torch.cat([d00, layer3], 1))
    d10 = _cbr(sd, "de_conv1_0", up(d01, layer2))
    d11 = _cbr(sd, "de_conv1_1", torch.cat([d10, layer2], 1))
    d20 = _cbr(sd, "de_conv2_0", up(d11, layer1))
    d21 = _cbr(sd, "de_conv2_1", torch.cat([d20, layer1], 1))
    d30 = _cbr(sd, "de_conv3_0", up(d21, conv1))
    d31 = _cbr(sd, "de_conv3_1", torch.cat([d30, conv1], 1))
    d40 = _cbr(sd, "de_conv4_0", F.interpolate(d31, patches.shape[-2:], mode="bilinear"))
    tr.update(de_conv0_1=d01, de_conv1_1=d11, de_conv2_1=d21, de_conv3_1=d31, de_conv4_0=d40)
    pred = _conv(sd, "pred", d40, 1, 1)
    weight = _conv(sd, "weight_pred", d40, 1, 1)
    tr.update(pred_raw=pred, weight_raw=weight)
    return pred, weight, d40


def _merge(pred_raw, weight_raw, bs, confidence, fov, nrows, patch, erp_size):
    """Heads + confidence-weighted ERP blend (spherical_model_iterative.py:371-380)."""
    pred = F.relu(pred_raw)
    if confidence:
        w = torch.sigmoid(weight_raw)
        pred = pred * w
        W = pers2equi(_unfold(w, bs), fov, nrows, patch, erp_size, "weight")
        zero = (W <= 1e-8).type(torch.float32)
        D = pers2equi(_unfold(pred, bs), fov, nrows, patch, erp_size, "pred")
        return D / (W + 1e-8 * zero)
    return pers2equi(_unfold(pred, bs), fov, nrows, patch, erp_size, "pred")


@torch.no_grad()
def forward_iterative(sd, high_res, iters, confidence=False, nrows=4, fov=(80, 80),
                      patch_size=(128, 128), trace=None):
    """spherical_fusion.forward(high_res, iter, confidence) -> list of `iters`
    tensors (B,1,He,We).  spherical_model_iterative.py:308-456."""
    sd = strip_module_prefix(sd)
    bs = high_res.shape[0]
    erp_size = tuple(high_res.shape[-2:])
    ph, pw = _pair(patch_size)
    low = (ph // 4, pw // 4)
    tr = trace if trace is not None else {}

    patches, _, _, _ = equi2pers(high_res, fov, nrows, (ph, pw))
    _, xyz, _, _ = equi2pers(high_res, fov, nrows, low)
    n_patch = patches.shape[-1]
    pf = _mlp_points(sd, "mlp_points1", xyz.contiguous())          # (N,64,p,p), batch-independent
    pf = pf.unsqueeze(0).expand(bs, -1, -1, -1, -1).reshape(bs * n_patch, *pf.shape[1:])
    t0 = {}
    pred, weight, _ = patch_network(sd, _fold(patches), pf, bs, trace=t0)
    tr["iter0"] = t0
    preds = [_merge(pred, weight, bs, confidence, fov, nrows, (ph, pw), erp_size)]

    for i in range(iters - 1):
        depth_p, _, _, _ = equi2pers(preds[i], fov, nrows, low)    # (B,1,p,p,N)
        # xyz (N,3,p,p) scaled by the previous depth (:388-391)
        pts = xyz.unsqueeze(0) * _fold(depth_p).reshape(bs, n_patch, 1, *low)
        pf = _mlp_points(sd, "mlp_points2", pts.reshape(bs * n_patch, 3, *low))
        ti = {}
        pred, weight, _ = patch_network(sd, _fold(patches), pf, bs, trace=ti)
        tr[f"iter{i + 1}"] = ti
        preds.append(_merge(pred, weight, bs, confidence, fov, nrows, (ph, pw), erp_size))
    return preds


@torch.no_grad()
def forward_single(sd, rgb, confidence=True, nrows=4, fov=(80, 80), patch_size=(128, 128), trace=None):
    """Single-stage spherical_fusion.forward(rgb, confidence) -> (B,1,He,We).
    spherical_model.py:238-314; the point MLP sees [cx, cy, 1, cx, cy] (:245-252)."""
    sd = strip_module_prefix(sd)
    bs = rgb.shape[0]
    erp_size = tuple(rgb.shape[-2:])
    ph, pw = _pair(patch_size)
    low = (ph // 4, pw // 4)
    patches, _, _, _ = equi2pers(rgb, fov, nrows, (ph, pw))
    _, _, uv, center_p = equi2pers(rgb, fov, nrows, low)
    n_patch = patches.shape[-1]
    cp = center_p.reshape(-1, 2, 1, 1).repeat(1, 1, *low)
    rho = torch.ones((uv.shape[0], 1, *low), dtype=torch.float32)
    pf = _mlp_points(sd, "mlp_points", torch.cat([cp, rho, cp], 1).contiguous())
    pf = pf.unsqueeze(0).expand(bs, -1, -1, -1, -1).reshape(bs * n_patch, *pf.shape[1:])
    tr = trace if trace is not None else {}
    pred, weight, _ = patch_network(sd, _fold(patches), pf, bs, down_name="down", trace=tr)
    return _merge(pred, weight, bs, confidence, fov, nrows, (ph, pw), erp_size)


@torch.no_grad()
def forward_360d(sd, high_res, fov, patch_size, nrows):
    """network_360d.py:308-380: encoder/decoder without point features and without the transformer, one pass, plain
    pers2equi blend of relu(pred)."""
    sd = strip_module_prefix(sd)
    bs = high_res.shape[0]
    erp_size = tuple(high_res.shape[-2:])
    ph, pw = _pair(patch_size)
    patches, _, _, _ = equi2pers(high_res, fov, nrows, (ph, pw))
    pred, weight, _ = patch_network(sd, _fold(patches), None, bs, use_points=False, use_transformer=False)
    return _merge(pred, weight, bs, False, fov, nrows, (ph, pw), erp_size)


@torch.no_grad()
def forward_test(sd, high_res, fov, patch_size, nrows, iters):
    """network_test.py:308-460: the iterative network with down1 512 -> 8 (256x256 patches) and the plain blend.
    Like the reference, iterations >= 1 return the PATCH prediction relu(pred) (B,1,P,P,N) - the pers2equi of the
    refinement pass is commented out there (:441-445) - so only iters <= 2 is meaningful."""
    assert iters in (1, 2)
    tr = {}
    outs = forward_iterative(sd, high_res, iters, False, nrows=nrows, fov=fov, patch_size=_pair(patch_size), trace=tr)
    if iters == 2:
        outs = [outs[0], _unfold(torch.relu(tr["iter1"]["pred_raw"]), high_res.shape[0])]
    return outs


def abs_rel_error(pred, gt, mask):
    """metrics.py:7-9."""
    return ((pred[mask > 0] - gt[mask > 0]).abs() / gt[mask > 0]).mean()


def eval_metrics(pred, gt, mask, median_scale=True):
    """test.py:151-170 + metrics.py:7-26 -> dict of the 7 metrics and the valid count."""
    pred = pred.clone()
    if median_scale:
        pred = pred * (gt[mask > 0].median() / pred[mask > 0].median())
    m = mask > 0
    p, g = pred[m], gt[m]
    ml = m & (pred > 1e-7) & (gt > 1e-7)
    ratio = torch.max(p / g, g / p)
    return {
        "abs_rel": ((p - g).abs() / g).mean().item(),
        "sq_rel": (((p - g) ** 2) / g).mean().item(),
        "rms_sq_lin": ((p - g) ** 2).mean().item(),
        "rms_sq_log": ((pred[ml].log() - gt[ml].log()) ** 2).mean().item(),
        "d1": (ratio < 1.25).float().mean().item(),
        "d2": (ratio < 1.25 ** 2).float().mean().item(),
        "d3": (ratio < 1.25 ** 3).float().mean().item(),
        "n": int(mask.sum().item()),
    }
