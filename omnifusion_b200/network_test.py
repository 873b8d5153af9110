"""The 256x256-patch variant of the iterative network behind the reference's interface (network_test.py:253-460):
``down1`` is 512 -> 8 over the 8x8 layer4 map (:271; token width 8*8*8 = 512), geometry is a call-time argument,
and the ERP merge is the plain pers2equi blend (the confidence branch is commented out there: :373-378, :442-447).
``spherical_fusion().forward(high_res, fov, patch_size, nrows, iter, confidence=True) -> [ERP depth (B,1,He,We)]`` for
iter = 1 and ``[ERP depth, patch prediction (B,1,P,P,N)]`` for iter = 2: the reference appends the un-merged patch
prediction for refinement passes (its pers2equi is commented out, :441-445) and cannot run iter > 2.
pos_emb is built for 18 patches (:273), so nrows must be 4, as in the reference."""
from . import tables
from .model._fusion import SphericalFusionBase


class spherical_fusion(SphericalFusionBase):
    KIND = "test"
    PATCH_SIZES = ((256, 256),)

    def __init__(self):
        super().__init__(4, 18, (256, 256), (80, 80))

    def forward(self, high_res, fov, patch_size, nrows, iter, confidence=True):
        self.fov, self.nrows = tables.pair(fov), nrows
        if tables.pair(patch_size) != (256, 256):
            raise ValueError("network_test's down1 (512 -> 8) fits 256x256 patches only (token width 512)")
        if tables.NUM_PATCHES.get(nrows) != 18:
            raise ValueError("network_test builds pos_emb for 18 patches: nrows must be 4")
        if iter not in (1, 2):
            raise ValueError("network_test returns patch predictions after the first pass: iter must be 1 or 2")
        if iter == 2 and high_res.shape[0] > 32:
            raise ValueError("iter=2 returns the patch prediction of the engine's last chunk: at most 32 panoramas per call")
        outs = self._run(high_res, iter, False)
        if iter == 2:
            bs = high_res.shape[0]
            pred = self.activation("pred_patch")                       # (B*N, P, P, 1) of the last pass
            outs = [outs[0], pred.reshape(bs, 18, 256, 256, 1).permute(0, 4, 2, 3, 1).contiguous()]
        return outs
