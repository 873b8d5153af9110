#!/usr/bin/env python
"""One eager forward of the benchmark configuration (for ncu captures; not a benchmark)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omnifusion_b200.checkpoint import synthetic_state_dict
from omnifusion_b200.model.spherical_model_iterative import spherical_fusion
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
net = spherical_fusion(4, 18, (128, 128), (80, 80))
net.load_state_dict(synthetic_state_dict("iterative", 18, 0))
net = net.to("cuda:0").eval()
x = torch.rand(B, 3, 512, 1024, generator=torch.Generator().manual_seed(123)).to("cuda:0")
with torch.no_grad():
    for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
        out = net(x, iter=2, confidence=True)
torch.cuda.synchronize()
print("depth mean", float(out[-1].mean()))
