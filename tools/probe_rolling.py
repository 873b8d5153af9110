#!/usr/bin/env python
"""Clock stamps of epilogue warp 0 of CTA 0 for the two rolling-row kernels (de_conv4_0 with the fused upsample, the
heads) at the benchmark size: per output row wait | ldtm(+zero)+arrive | math | wait_group | sts | fence | store, and the
row period.  Timing experiment."""
import ctypes as C
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from omnifusion_b200 import _lib
import ofb_ops as o

DEV = torch.device("cuda:0")
L = _lib.lib()
n = 144
x = o.split16(torch.randn(n, 64, 64, 32, device=DEV))
w = torch.randn(32, 3, 3, 32, device=DEV) * (1.0 / (32 * 9)) ** 0.5
mul = o.weight_scale(w)
ws = o.split16(w, mul)
out = torch.empty(2 * n * 128 * 128 * 32, dtype=torch.float16, device=DEV)
scale = torch.ones(32, device=DEV); shift = torch.zeros(32, device=DEV)
d = _lib.ConvDesc()
d.in0 = x.data_ptr(); d.c0 = 32; d.c1 = 0; d.n, d.h, d.w = n, 128, 128
d.wgt = w.data_ptr(); d.k, d.stride, d.pad, d.cout = 3, 1, 1, 32
d.scale, d.shift = scale.data_ptr(), shift.data_ptr()
d.act, d.out, d.engine, d.in_fmt, d.out_fmt = 1, out.data_ptr(), _lib.ENGINE_TC, 1, 1
d.wgt_split, d.wgt_unscale = ws.data_ptr(), 1.0 / mul
d.ups2x = 1
st = _lib.stream_of(DEV)


def stamps(fn, name):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    _lib.check(L.ofb_debug_set(16))
    fn(); fn()
    buf = np.zeros((512, 8), dtype=np.int64)
    _lib.check(L.ofb_debug_stamps(buf.ctypes.data))
    _lib.check(L.ofb_debug_set(0))
    T = buf[20:60]
    per = np.diff(T[:, 0])
    ph = np.diff(T[:-1], axis=1)
    print(f"{name}: {us:.1f} us per launch; row period median {np.median(per):.0f} clk; phases median "
          f"(wait | ldtm+arrive | math | wait_group | sts | fence | store) = {[int(v) for v in np.median(ph, axis=0)]}")


xh = o.split16(torch.relu(torch.randn(n, 128, 128, 32, device=DEV)))
w16 = torch.zeros(16, 3, 3, 32, device=DEV); w16[:2] = torch.randn(2, 3, 3, 32, device=DEV) * 0.1
mul16 = o.weight_scale(w16)
ws16 = o.split16(w16, mul16)
pairs = torch.empty(n, 128, 128, 2, device=DEV)
for ns in (2, 0):
    os.environ["OFB_NSTACK"] = str(ns)
    _lib.check(L.ofb_debug_nstack(ns))
    stamps(lambda: _lib.check(L.ofb_conv_f32(C.byref(d), st)), f"de_conv4_0 fused upsample nstack={ns}")
    stamps(lambda: _lib.check(L.ofb_heads_tc_pairs_f16(_lib.ptr(xh), n, 128, 128, _lib.ptr(ws16), 1.0 / mul16, 0.5, -0.5,
                                                        _lib.ptr(pairs), st)), f"heads nstack={ns}")

# MMA warp stamps of the tap-stacked scheme ("tc_debug" & 2048): per input row, clocks waiting for the ring row
# (producers), claiming accumulator blocks (epilogue) and issuing the MMAs
_lib.check(L.ofb_debug_nstack(2))
for name, fn in (("de_conv4_0 fused upsample", lambda: _lib.check(L.ofb_conv_f32(C.byref(d), st))),
                 ("heads", lambda: _lib.check(L.ofb_heads_tc_pairs_f16(_lib.ptr(xh), n, 128, 128, _lib.ptr(ws16), 1.0 / mul16, 0.5, -0.5,
                                                                       _lib.ptr(pairs), st)))):
    _lib.check(L.ofb_debug_set(2048))
    fn(); fn()
    buf = np.zeros((512, 8), dtype=np.int64)
    _lib.check(L.ofb_debug_stamps(buf.ctypes.data))
    _lib.check(L.ofb_debug_set(0))
    T = buf[20:100]
    ok = T[:, 3] > 0
    T = T[ok]
    print(f"{name}: MMA warp per input row (median clk): wait ring row {np.median(T[:,1]-T[:,0]):.0f} | claim blocks "
          f"{np.median(np.where(T[:,2] > 0, T[:,2]-T[:,1], 0)):.0f} | issue + commit {np.median(np.where(T[:,2] > 0, T[:,3]-T[:,2], T[:,3]-T[:,1])):.0f} | period {np.median(np.diff(T[:,0])):.0f}")
