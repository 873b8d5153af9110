#!/usr/bin/env python
"""%globaltimer stamps of CTA 0 inside the LAST image-stationary chain launch of one graph replay ("tc_debug" & 256):
per (layer, image slot) dependency release, first operands, last MMA issued, accumulator complete, stores issued,
stores complete.  Timing experiment."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omnifusion_b200 import _lib
from omnifusion_b200.checkpoint import synthetic_state_dict
from omnifusion_b200.model.spherical_model_iterative import spherical_fusion

chain = int(sys.argv[1]) if len(sys.argv) > 1 else 2
net = spherical_fusion(4, 18, (128, 128), (80, 80))
net.load_state_dict(synthetic_state_dict("iterative", 18, 0))
net = net.to("cuda:0").eval()
net.set_option("tc_debug", 256)
net.set_option("chain", chain)
x = torch.rand(8, 3, 512, 1024, generator=torch.Generator().manual_seed(123)).to("cuda:0")
with torch.no_grad():
    for _ in range(3):
        net.forward_graphed(x, 2, True)
    torch.cuda.synchronize()
    L = _lib.lib()
    L.ofb_debug_timeline(None, 0)
    net.forward_graphed(x, 2, True)
    torch.cuda.synchronize()
    buf = (C.c_longlong * (1024 * 8))()
    L.ofb_debug_timeline_raw(buf)
rows = [[buf[(512 + i) * 8 + k] for k in range(6)] for i in range(96)]
rows = [(i, r) for i, r in enumerate(rows) if any(r)]
t0 = min(v for _, r in rows for v in r if v)
print("layer slot | dep_released first_ops last_mma_issued acc_complete stores_issued stores_complete (us from the first stamp)")
for i, r in rows:
    print(f"{i // 8:5d} {i % 8:4d} | " + " ".join(f"{(v - t0) / 1e3:9.2f}" if v else "        -" for v in r))
