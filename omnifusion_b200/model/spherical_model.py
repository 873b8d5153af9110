"""Single-stage OmniFusion network behind the reference's interface
(model/spherical_model.py:190-314): ``spherical_fusion(...).forward(rgb, confidence=True)
-> (B,1,He,We)``.  Its point MLP sees the constant [cx, cy, 1, cx, cy] per patch (:245-252)."""
import torch

from ._fusion import SphericalFusionBase


class spherical_fusion(SphericalFusionBase):
    KIND = "single"

    def _point_table(self, low_geo, device):
        n = low_geo["grid"].shape[0]
        p = low_geo["grid"].shape[1]
        cp = low_geo["center_p"].to(device).reshape(n, 2, 1, 1).repeat(1, 1, p, p)
        rho = torch.ones((n, 1, p, p), dtype=torch.float32, device=device)
        return torch.cat([cp, rho, cp], 1)

    def forward(self, rgb, confidence=True):
        return self._run(rgb, 1, confidence)[0]
