"""Test-side convenience wrappers: call libofb's C ABI operators on torch CUDA tensors."""
import ctypes as C

import torch

from omnifusion_b200 import _lib

L = _lib.lib
ck = _lib.check
p = _lib.ptr


def _st(t):
    _lib.use_device(t.device)
    return _lib.stream_of(t.device)


def nhwc(x):
    """(n,c,h,w) -> contiguous (n,h,w,c)"""
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def ohwi(w):
    """(O,I,kh,kw) -> (O,kh,kw,I)"""
    return w.permute(0, 2, 3, 1).contiguous()


def conv(in0, wgt, k, stride, pad, in1=None, scale=None, shift=None, residual=None, act=0, engine=0):
    n, h, w, c0 = in0.shape
    cout = wgt.shape[0]
    oh = (h + 2 * pad - k) // stride + 1
    ow = (w + 2 * pad - k) // stride + 1
    out = torch.empty((n, oh, ow, cout), device=in0.device, dtype=torch.float32)
    d = _lib.ConvDesc()
    d.in0, d.in1 = in0.data_ptr(), (in1.data_ptr() if in1 is not None else None)
    d.c0, d.c1 = c0, (in1.shape[3] if in1 is not None else 0)
    d.n, d.h, d.w = n, h, w
    d.wgt, d.k, d.stride, d.pad, d.cout = wgt.data_ptr(), k, stride, pad, cout
    d.scale = scale.data_ptr() if scale is not None else None
    d.shift = shift.data_ptr() if shift is not None else None
    d.residual = residual.data_ptr() if residual is not None else None
    d.act, d.out, d.engine = act, out.data_ptr(), engine
    ck(L().ofb_conv_f32(C.byref(d), _st(in0)))
    return out


def equi2pers(erp, grid, layout):
    b, c, he, we = erp.shape
    n, ph, pw, _ = grid.shape
    if layout == _lib.LAYOUT_REF:
        out = torch.empty((b, c, ph, pw, n), device=erp.device)
    else:
        out = torch.empty((b * n, ph, pw, 4 if c == 3 else c), device=erp.device)
    ck(L().ofb_equi2pers_f32(p(erp), b, c, he, we, p(grid), n, ph, pw, p(out), layout, _st(erp)))
    return out


def equi2pers_taps(grid, he, we):
    n, ph, pw, _ = grid.shape
    x0 = torch.empty((n, ph, pw), device=grid.device, dtype=torch.int32)
    y0 = torch.empty_like(x0)
    ck(L().ofb_equi2pers_taps(p(grid), n, ph, pw, he, we, p(x0), p(y0), _st(grid)))
    return x0, y0


def pers2equi(pers, tab, he, we, layout, dims=None):
    if layout == _lib.LAYOUT_REF:
        b, c, ph, pw, n = pers.shape
    else:
        b, c, n, ph, pw = dims
    out = torch.empty((b, c, he, we), device=pers.device)
    ck(L().ofb_pers2equi_f32(p(pers), b, c, n, ph, pw, layout, p(tab["rowptr"]), p(tab["idx"]), p(tab["w"]),
                             he, we, p(out), _st(pers)))
    return out


def blend_conf(pred_w, conf, b, n, tab, he, we):
    ph, pw = pred_w.shape[-2:]
    out = torch.empty((b, 1, he, we), device=pred_w.device)
    ck(L().ofb_blend_conf_f32(p(pred_w), p(conf), b, n, ph, pw, p(tab["rowptr"]), p(tab["idx"]), p(tab["w"]),
                              he, we, p(out), _st(pred_w)))
    return out


def stem(x, wgt, scale, shift):
    n, h, w, _ = x.shape
    out = torch.empty((n, h // 2, w // 2, 64), device=x.device)
    ck(L().ofb_stem_f32(p(x), n, h, w, p(wgt), p(scale), p(shift), p(out), 0, _st(x)))
    return out


def maxpool(x):
    n, h, w, c = x.shape
    out = torch.empty((n, h // 2, w // 2, c), device=x.device)
    ck(L().ofb_maxpool3x3s2_f32(p(x), n, h, w, c, p(out), 0, _st(x)))
    return out


def upsample2x(x, img_bias=None):
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), device=x.device)
    ck(L().ofb_upsample2x_f32(p(x), p(img_bias), n, h, w, c, p(out), 0, _st(x)))
    return out


def point_embed(pts, depth, imgs, w1, s1, t1, w2, s2, t2, base):
    n, cin, pp, _ = pts.shape
    out = torch.empty((imgs, pp, pp, 64), device=pts.device)
    ck(L().ofb_point_embed_f32(p(pts), n, cin, pp, p(depth), imgs, p(w1), p(s1), p(t1), p(w2), p(s2), p(t2),
                               p(base), p(out), 0, _st(pts)))
    return out


def token_pack(down, pos, n_patch):
    imgs = down.shape[0]
    out = torch.empty((imgs, 512), device=down.device)
    ck(L().ofb_token_pack_f32(p(down), p(pos), imgs, n_patch, p(out), 0, _st(down)))
    return out


def layernorm(x, g, b, eps):
    out = torch.empty_like(x)
    ck(L().ofb_layernorm_f32(p(x), p(g), p(b), x.shape[0], x.shape[1], eps, p(out), 0, 0, _st(x)))
    return out


def attention(q, kv, bs, n, heads=4, hd=128):
    out = torch.empty_like(q)
    ck(L().ofb_attention_f32(p(q), p(kv), bs, n, heads, hd, p(out), 0, _st(q)))
    return out


def heads(x, wp, bp, wc, bc, confidence):
    imgs, h, w, _ = x.shape
    pred = torch.empty((imgs, h, w), device=x.device)
    conf = torch.empty((imgs, h, w), device=x.device)
    ck(L().ofb_heads_f32(p(x), imgs, h, w, p(wp), bp, p(wc), bc, int(confidence), p(pred), p(conf), 0, _st(x)))
    return pred, conf


def absrel(pred, gt, mask, scale=1.0):
    out = torch.zeros(2, dtype=torch.float64, device=pred.device)
    m = mask.to(torch.uint8).contiguous()
    ck(L().ofb_absrel_partial(p(pred), p(gt), p(m), pred.numel(), scale, p(out), _st(pred)))
    return out


def split16(x, mul=1.0):
    """float32 tensor -> split-half planes (2*numel halves: hi plane then lo plane)."""
    x = x.contiguous()
    out = torch.empty(2 * x.numel(), dtype=torch.float16, device=x.device)
    ck(L().ofb_split_f16(p(x), x.numel(), mul, p(out), _st(x)))
    return out


def merge16(planes, shape):
    out = torch.empty(shape, dtype=torch.float32, device=planes.device)
    ck(L().ofb_merge_f16(p(planes), out.numel(), p(out), _st(planes)))
    return out


def weight_scale(w):
    """power-of-two multiplier that puts max|w| in [2^13, 2^14) for the fp16 hi/lo planes"""
    import math
    m = float(w.abs().max())
    return 2.0 ** (13 - math.floor(math.log2(m))) if m > 0 else 1.0


def conv_fmt(in0, wgt, k, stride, pad, in1=None, scale=None, shift=None, residual=None, act=0, engine=0,
             in_fmt=0, out_fmt=0, ups2x=0, ksplit=1):
    """conv with explicit storage formats: float32 NHWC tensors in, converted to/from split planes here.
    ups2x: in0 is the low-resolution tensor, bilinearly upsampled x2 inside the conv kernel."""
    n, h, w, c0 = in0.shape
    if ups2x:
        h, w = 2 * h, 2 * w
    cout = wgt.shape[0]
    oh = (h + 2 * pad - k) // stride + 1
    ow = (w + 2 * pad - k) // stride + 1
    keep = []
    def cvt(t, fmt):
        if t is None:
            return None
        t = split16(t) if fmt == 1 else t
        keep.append(t)
        return t.data_ptr()
    d = _lib.ConvDesc()
    d.in0, d.in1 = cvt(in0, in_fmt), cvt(in1, in_fmt)
    d.c0, d.c1 = c0, (in1.shape[3] if in1 is not None else 0)
    d.n, d.h, d.w = n, h, w
    d.wgt, d.k, d.stride, d.pad, d.cout = wgt.data_ptr(), k, stride, pad, cout
    mul = weight_scale(wgt)
    ws = split16(wgt, mul)
    d.wgt_split, d.wgt_unscale = ws.data_ptr(), 1.0 / mul
    d.scale = scale.data_ptr() if scale is not None else None
    d.shift = shift.data_ptr() if shift is not None else None
    d.residual = cvt(residual, out_fmt)
    numel = n * oh * ow * cout
    out = torch.empty(2 * numel, dtype=torch.float16, device=in0.device) if out_fmt == 1 else \
        torch.empty((n, oh, ow, cout), device=in0.device)
    d.act, d.out, d.engine, d.in_fmt, d.out_fmt = act, out.data_ptr(), engine, in_fmt, out_fmt
    d.ups2x = ups2x
    if ksplit > 1:
        # split-K: the kernel leaves raw float32 partial sums (ksplit, n*oh*ow, cout); returns them with 1/mul
        part = torch.zeros(ksplit, n * oh * ow, cout, device=in0.device)
        d.ksplit, d.partial = ksplit, part.data_ptr()
        ck(L().ofb_conv_f32(C.byref(d), _st(in0)))
        return part, 1.0 / mul
    ck(L().ofb_conv_f32(C.byref(d), _st(in0)))
    return merge16(out, (n, oh, ow, cout)) if out_fmt == 1 else out


def last_conv_variant():
    return L().ofb_last_conv_variant().decode()
