#!/usr/bin/env python
"""Timing experiments on the tcgen05 conv engine (not a test, not a benchmark line): runs one layer shape
with the CTA pairing and parts of the epilogue switched off (TcParams::dbg) to see what bounds it."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from omnifusion_b200 import _lib
import ofb_ops as o

DEV = torch.device("cuda:0")
L = _lib.lib()
h = C.c_void_p()
_lib.check(L.ofb_create(0, C.byref(h)))


def opt(k, v):
    _lib.check(L.ofb_set_option(h, k.encode(), int(v)))


def make(n, hw, cin, cout, k=3, stride=1):
    x = o.split16(torch.randn(n, hw, hw, cin, device=DEV))
    w = torch.randn(cout, k, k, cin, device=DEV) * (1.0 / (cin * k * k)) ** 0.5
    mul = o.weight_scale(w)
    ws = o.split16(w, mul)
    oh = hw // stride
    out = torch.empty(2 * n * oh * oh * cout, dtype=torch.float16, device=DEV)
    scale = torch.ones(cout, device=DEV); shift = torch.zeros(cout, device=DEV)
    d = _lib.ConvDesc()
    d.in0 = x.data_ptr(); d.c0 = cin; d.c1 = 0; d.n, d.h, d.w = n, hw, hw
    d.wgt = w.data_ptr(); d.k, d.stride, d.pad, d.cout = k, stride, k // 2, cout
    d.scale, d.shift = scale.data_ptr(), shift.data_ptr()
    d.act, d.out, d.engine, d.in_fmt, d.out_fmt = 1, out.data_ptr(), _lib.ENGINE_TC, 1, 1
    d.wgt_split, d.wgt_unscale = ws.data_ptr(), 1.0 / mul
    return d, (x, w, ws, out, scale, shift)


def time_conv(d, iters=30):
    st = _lib.stream_of(DEV)
    for _ in range(5):
        _lib.check(L.ofb_conv_f32(C.byref(d), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        _lib.check(L.ofb_conv_f32(C.byref(d), st))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


shapes = [(144, 8, 256, 256), (144, 16, 128, 128), (144, 4, 512, 512), (576, 8, 256, 256)]
for (n, hw, cin, cout) in shapes:
    d, keep = make(n, hw, cin, cout)
    fl = 2.0 * n * hw * hw * cout * 9 * cin
    row = []
    for cta2 in (0, 1):
        opt("cta2", cta2)
        for dbg, nm in ((0, "full"), (64, "no epilogue math/stores"), (128, "no global stores")):
            opt("tc_debug", dbg)
            us = time_conv(d)
            row.append(f"cta2={cta2} {nm}={us:.1f}us ({fl / us / 1e6:.0f} TFLOP/s)" if dbg == 0 else f"cta2={cta2} {nm}={us:.1f}us")
    opt("tc_debug", 0)
    opt("cta2", 1)
    print(f"n={n} {hw}x{hw} c{cin}->o{cout} ({fl / 1e9:.1f} GFLOP): " + "  ".join(row), flush=True)
