#!/usr/bin/env python
"""Clock stamps of one epilogue warp of CTA 0 for the mid-layer (cta_group::2) conv shapes at the benchmark batch
("tc_debug" & 16): where the un-overlapped epilogue of a one-tile-per-CTA launch spends its time.  Timing
experiment, not a benchmark."""
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


def make(n, hw, cin, cout, res):
    x0 = o.split16(torch.randn(n, hw, hw, cin, device=DEV))
    w = torch.randn(cout, 3, 3, cin, device=DEV) * (1.0 / (cin * 9)) ** 0.5
    mul = o.weight_scale(w)
    ws = o.split16(w, mul)
    out = torch.empty(2 * n * hw * hw * cout, dtype=torch.float16, device=DEV)
    r = o.split16(torch.randn(n, hw, hw, cout, device=DEV)) if res else None
    scale = torch.ones(cout, device=DEV); shift = torch.zeros(cout, device=DEV)
    d = _lib.ConvDesc()
    d.in0 = x0.data_ptr(); d.c0 = cin; d.c1 = 0; d.n, d.h, d.w = n, hw, hw
    d.wgt = w.data_ptr(); d.k, d.stride, d.pad, d.cout = 3, 1, 1, cout
    d.scale, d.shift = scale.data_ptr(), shift.data_ptr()
    d.residual = r.data_ptr() if res else None
    d.act, d.out, d.engine, d.in_fmt, d.out_fmt = 1, out.data_ptr(), _lib.ENGINE_TC, 1, 1
    d.wgt_split, d.wgt_unscale = ws.data_ptr(), 1.0 / mul
    return d, (x0, w, ws, out, scale, shift, r)


# a handle only to set the per-handle debug option; ofb_conv_f32 itself uses the default options, so run the conv
# through a forward-free path: the option scope is installed by ofb_forward_f32 only -> use the env override instead
h = C.c_void_p()
_lib.check(L.ofb_create(0, C.byref(h)))
for (nm, n, hw, cin, cout, res) in [("layer3 c256@8 +res", 144, 8, 256, 256, True), ("layer2 c128@16", 144, 16, 128, 128, False),
                                    ("layer4 c512@4 +res", 144, 4, 512, 512, True)]:
    d, keep = make(n, hw, cin, cout, res)
    st = _lib.stream_of(DEV)
    _lib.check(L.ofb_debug_set(16))
    for _ in range(3):
        _lib.check(L.ofb_conv_f32(C.byref(d), st))
    buf = np.zeros((512, 8), dtype=np.int64)
    _lib.check(L.ofb_debug_stamps(buf.ctypes.data))
    _lib.check(L.ofb_debug_set(0))
    print(nm, "variant", o.last_conv_variant(), "stamps per chunk of epilogue warp 0 (clk): wait_tfull | ldtm+arrive | math | wait_group | sts | fence | store")
    for k in range(2):
        r = buf[k]
        print("   tile", k, [int(r[j + 1] - r[j]) for j in range(7)])
