#!/usr/bin/env python
"""Golden rays for omnifusion_b200.pointcloud from the reference's own coords2uv / uv2xyz (util.py:159-174),
extracted by source (util.py imports matplotlib / OpenEXR, absent here) and run on the meshgrid of
test.py:210-213.  Run in the authoring container: python tests/golden/make_golden_pointcloud.py"""
import ast
import os
import numpy as np

src = open("/root/reference/util.py").read()
mod = ast.parse(src)
ns = {"np": np}
for node in mod.body:
    if isinstance(node, ast.FunctionDef) and node.name in ("coords2uv", "uv2xyz"):
        exec(compile(ast.Module([node], []), "util.py", "exec"), ns)
out = {}
for (h, w) in [(8, 16), (64, 128)]:
    coords = np.stack(np.meshgrid(range(w), range(h)), -1)
    coords = np.reshape(coords, [-1, 2])
    coords += 1
    uv = ns["coords2uv"](coords, w, h)
    out[f"rays_{h}x{w}"] = ns["uv2xyz"](uv)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pointcloud_rays.npz"), **out)
print({k: v.shape for k, v in out.items()})
