"""Iterative (coarse -> refine) OmniFusion network behind the reference's interface
(model/spherical_model_iterative.py:253-456): ``spherical_fusion(nrows, npatches, patch_size,
fov).forward(high_res, iter, confidence=False) -> list of `iter` tensors (B,1,He,We)``."""
from ._fusion import SphericalFusionBase


class spherical_fusion(SphericalFusionBase):
    KIND = "iterative"

    def __init__(self, nrows=4, npatches=18, patch_size=(128, 128), fov=(80, 80)):
        # the reference's default patch_size=(256,256) (:254) shape-errors in its own forward;
        # both of its trainers pass (128,128) (train_erp_depth_iterative.py:46,142)
        super().__init__(nrows, npatches, patch_size, fov)

    def forward(self, high_res, iter, confidence=False):
        return self._run(high_res, iter, confidence)
