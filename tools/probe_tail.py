#!/usr/bin/env python
"""Timing experiments on the narrow (HBM / fabric bound) decoder-tail layers: each shape is run with parts of
the kernel disabled (TcParams::dbg; results are wrong in those runs) to see which stage bounds it."""
import ctypes as C
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from omnifusion_b200 import _lib
import ofb_ops as o

DEV = torch.device("cuda:0")
L = _lib.lib()
h = C.c_void_p()
_lib.check(L.ofb_create(0, C.byref(h)))


def opt(k, v):
    _lib.check(L.ofb_set_option(h, k.encode(), int(v)))


def make(n, hw, c0, c1, cout, ups=0):
    ih = hw // 2 if ups else hw
    x0 = o.split16(torch.randn(n, ih, ih, c0, device=DEV))
    x1 = o.split16(torch.randn(n, ih, ih, c1, device=DEV)) if c1 else None
    cin = c0 + c1
    w = torch.randn(cout, 3, 3, cin, device=DEV) * (1.0 / (cin * 9)) ** 0.5
    mul = o.weight_scale(w)
    ws = o.split16(w, mul)
    out = torch.empty(2 * n * hw * hw * cout, dtype=torch.float16, device=DEV)
    scale = torch.ones(cout, device=DEV); shift = torch.zeros(cout, device=DEV)
    d = _lib.ConvDesc()
    d.in0 = x0.data_ptr(); d.c0 = c0; d.c1 = c1; d.n, d.h, d.w = n, hw, hw
    if c1: d.in1 = x1.data_ptr()
    d.wgt = w.data_ptr(); d.k, d.stride, d.pad, d.cout = 3, 1, 1, cout
    d.scale, d.shift = scale.data_ptr(), shift.data_ptr()
    d.act, d.out, d.engine, d.in_fmt, d.out_fmt = 1, out.data_ptr(), _lib.ENGINE_TC, 1, 1
    d.wgt_split, d.wgt_unscale = ws.data_ptr(), 1.0 / mul
    d.ups2x = ups
    return d, (x0, x1, w, ws, out, scale, shift)


def time_conv(d, iters=20):
    st = _lib.stream_of(DEV)
    for _ in range(3):
        _lib.check(L.ofb_conv_f32(C.byref(d), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        _lib.check(L.ofb_conv_f32(C.byref(d), st))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


shapes = [("de_conv4_0 ups", 144, 128, 32, 0, 32, 1), ("de_conv4_0 plain", 144, 128, 32, 0, 32, 0),
          ("de_conv3_1", 144, 64, 64, 64, 32, 0), ("de_conv3_0", 144, 64, 64, 0, 64, 0),
          ("layer1", 144, 32, 64, 0, 64, 0), ("de_conv2_1", 144, 32, 64, 64, 64, 0)]
names = {128: "noStore", 160: "noStore+noInterp", 0: "full", 3: "noMMA", 32: "noInterp", 64: "noEpi", 67: "noMMA+noEpi", 99: "noMMA+noInterp+noEpi", 12: "noLoads", 79: "noLoads+noMMA+noEpi"}
for direct in (0, 1):
  opt("direct32", direct)
  print("direct32 =", direct)
  for (nm, n, hw, c0, c1, cout, ups) in shapes[:3]:
    d, keep = make(n, hw, c0, c1, cout, ups)
    row = []
    for dbg in (0, 64, 128, 32, 160):
        opt("tc_debug", dbg)
        row.append(f"{names[dbg]}={time_conv(d):.1f}us")
    opt("tc_debug", 0)
    print(f"{nm} n={n} {hw}x{hw} c{c0}+{c1}->o{cout}: " + "  ".join(row), flush=True)
for (nm, n, hw, c0, c1, cout, ups) in []:
    d, keep = make(n, hw, c0, c1, cout, ups)
    row = []
    for dbg in ((0, 3, 32, 64, 67, 99) if ups else (0, 3, 64, 67, 12, 79)):
        opt("tc_debug", dbg)
        row.append(f"{names[dbg]}={time_conv(d):.1f}us")
    opt("tc_debug", 0)
    print(f"{nm} n={n} {hw}x{hw} c{c0}+{c1}->o{cout}: " + "  ".join(row), flush=True)

# ---- clock stamps of epilogue warp 2 of CTA 0
import numpy as np


opt("direct32", 0)
for (nm, n, hw, c0, c1, cout, ups) in shapes[:3]:
    d, keep = make(n, hw, c0, c1, cout, ups)
    opt("tc_debug", 16)
    st = _lib.stream_of(DEV)
    for _ in range(2):
        _lib.check(L.ofb_conv_f32(C.byref(d), st))
    buf = np.zeros((512, 8), dtype=np.int64)
    _lib.check(L.ofb_debug_stamps(buf.ctypes.data))
    opt("tc_debug", 0)
    T = buf[8:28]
    print(nm, "stamps per tile: wait_tfull | ldtm+arrive | math | wait_group | sts | fence | store ; period")
    for k in range(len(T) - 1):
        r = T[k]
        print("  ", [int(r[j + 1] - r[j]) for j in range(7)], int(T[k + 1][0] - r[0]))
