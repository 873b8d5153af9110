#!/usr/bin/env python
"""Per-launch timeline of the tcgen05 conv kernels inside one CUDA-graph replay of the benchmark forward
("tc_debug" & 256: %globaltimer stamps of CTA 0 of every launch).  Prints, per launch, the gap to the previous
kernel's end and the phases inside the kernel (ns).  Timing experiment, not a benchmark."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omnifusion_b200 import _lib
from omnifusion_b200.checkpoint import synthetic_state_dict
from omnifusion_b200.model.spherical_model_iterative import spherical_fusion

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
graph = (sys.argv[2] if len(sys.argv) > 2 else "graph") == "graph"
net = spherical_fusion(4, 18, (128, 128), (80, 80))
net.load_state_dict(synthetic_state_dict("iterative", 18, 0))
net = net.to("cuda:0").eval()
net.set_option("tc_debug", 256)
x = torch.rand(B, 3, 512, 1024, generator=torch.Generator().manual_seed(123)).to("cuda:0")
fwd = (lambda: net.forward_graphed(x, 2, True)) if graph else (lambda: net(x, iter=2, confidence=True))
with torch.no_grad():
    for _ in range(3):
        fwd()
    torch.cuda.synchronize()
    L = _lib.lib()
    L.ofb_debug_timeline(None, 0)            # reset
    fwd()
    torch.cuda.synchronize()
    buf = (C.c_longlong * (1024 * 8))()
    n = L.ofb_debug_timeline(buf, 1024)
rows = [[buf[i * 8 + k] for k in range(8)] for i in range(n)]
rows.sort(key=lambda r: r[0])
t0 = rows[0][0]
names = ["entry", "prolog", "dep_wait", "1st_ops", "mma_end", "acc_done", "st_issued", "end"]
print(f"{n} tcgen05 launches; columns: start(us) gap_from_prev_end | " + " ".join(f"+{m}" for m in names[1:]) + " (ns from entry)")
prev_end = None
tot_gap = tot_in = 0
for r in rows:
    gap = (r[0] - prev_end) if prev_end else 0
    rel = [(v - r[0]) if v else -1 for v in r[1:]]
    print(f"{(r[0] - t0) / 1e3:9.1f} {gap:7d} | " + " ".join(f"{v:7d}" for v in rel))
    prev_end = r[7] if r[7] else prev_end
    tot_gap += max(gap, 0); tot_in += (r[7] - r[0]) if r[7] else 0
print(f"sum of in-kernel time (CTA 0) {tot_in / 1e3:.1f} us, sum of gaps {tot_gap / 1e3:.1f} us, span {(rows[-1][7] - t0) / 1e3:.1f} us")
