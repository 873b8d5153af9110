"""The 360D ablation network behind the reference's interface (network_360d.py:253-380, driven by
test_360d_tmp.py:198): the same ResNet-34 patch encoder and decoder WITHOUT the point-feature add (:325) and
WITHOUT the transformer (:330-335), one pass, plain pers2equi blend of relu(pred) (no confidence: :371-379).
``spherical_fusion().forward(high_res, fov, patch_size, nrows, confidence=True) -> (B,1,He,We)``; geometry is a
call-time argument and `confidence` is accepted and ignored, exactly as in the reference.  The state_dict is the
iterative model's (the unused sub-modules still exist there)."""
from . import tables
from .model._fusion import SphericalFusionBase


class spherical_fusion(SphericalFusionBase):
    KIND = "iterative"

    def __init__(self):
        super().__init__(4, 18, (128, 128), (80, 80))
        self.set_option("no_point_feat", 1)
        self.set_option("no_transformer", 1)

    def forward(self, high_res, fov, patch_size, nrows, confidence=True):
        self.fov, self.nrows = tables.pair(fov), nrows
        self.patch_size = tables.pair(patch_size)
        if self.patch_size not in ((64, 64), (128, 128), (256, 256)):
            raise ValueError(f"patch_size must be (64,64), (128,128) or (256,256), got {self.patch_size}")
        self.npatches = tables.NUM_PATCHES[nrows] if nrows in tables.NUM_PATCHES else tables._centers(nrows)
        return self._run(high_res, 1, False)[0]
