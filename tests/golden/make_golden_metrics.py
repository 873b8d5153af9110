#!/usr/bin/env python
"""Golden values of the reference's metrics.py:7-26 (imported from /root/reference, it only needs torch) on seeded
inputs, for pinning oracle.model.eval_metrics.  Run in the authoring container:
python tests/golden/make_golden_metrics.py"""
import importlib.util
import json
import os
import torch

spec = importlib.util.spec_from_file_location("ref_metrics", "/root/reference/metrics.py")
rm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rm)


def urand(*shape, seed):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed))


out = {}
for tag, (seed_p, seed_g, scale_p) in {"a": (1, 2, 1.0), "b": (11, 12, 0.37)}.items():
    pred = (0.1 + 7.9 * urand(2, 1, 32, 64, seed=seed_p)) * scale_p
    gt = 0.1 + 9.0 * urand(2, 1, 32, 64, seed=seed_g)
    mask = ((gt <= 8) & (gt > 0.1)).to(torch.uint8)
    # test.py:161-162 median scaling over the whole batch tensor, then the seven metrics
    scaled = pred * (torch.median(gt[mask > 0]) / torch.median(pred[mask > 0]))
    for name, p in (("raw", pred), ("median_scaled", scaled)):
        out[f"{tag}_{name}"] = {
            "abs_rel": rm.abs_rel_error(p, gt, mask).item(), "sq_rel": rm.sq_rel_error(p, gt, mask).item(),
            "rms_sq_lin": rm.lin_rms_sq_error(p, gt, mask).item(), "rms_sq_log": rm.log_rms_sq_error(p, gt, mask).item(),
            "d1": rm.delta_inlier_ratio(p, gt, mask, 1).item(), "d2": rm.delta_inlier_ratio(p, gt, mask, 2).item(),
            "d3": rm.delta_inlier_ratio(p, gt, mask, 3).item(), "n": int(mask.sum().item()),
            "seeds": [seed_p, seed_g], "pred_scale": scale_p}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics_golden.json"), "w"), indent=1)
print(json.dumps(out["a_raw"]))
