"""Device side of the reference's Dataset.__getitem__ (dataset_loader_stanford.py:52,79): the decoded uint8 panorama
(cv2 channel order, HWC) becomes the float32 CHW network input, `rgb.astype(np.float32) / 255` then
`transpose(2, 0, 1)`.  Decoding and the INTER_AREA resize stay on the host (cv2); shipping the uint8 image and
converting on the GPU moves 4x fewer bytes over PCIe than the reference's float32 batch."""
import torch

from . import _lib


def rgb_u8_to_input(img_u8):
    """img_u8: (B,H,W,3) or (H,W,3) uint8 CUDA tensor -> (B,3,H,W) float32, bit-identical to the reference."""
    if img_u8.dim() == 3:
        img_u8 = img_u8.unsqueeze(0)
    if not img_u8.is_cuda:
        raise _lib.OfbError(f"img_u8 must be a CUDA tensor: omnifusion_b200 has no CPU path (got {img_u8.device})")
    if img_u8.dtype != torch.uint8 or img_u8.dim() != 4 or img_u8.shape[3] > 4:
        raise _lib.OfbError(f"img_u8 must be uint8 (B,H,W,C<=4), got {img_u8.dtype} {tuple(img_u8.shape)}")
    img_u8 = img_u8.contiguous()
    b, h, w, c = img_u8.shape
    out = torch.empty(b, c, h, w, dtype=torch.float32, device=img_u8.device)
    _lib.use_device(img_u8.device)
    _lib.check(_lib.lib().ofb_u8hwc_to_f32chw(_lib.ptr(img_u8), b, h, w, c, _lib.ptr(out), _lib.stream_of(img_u8.device)))
    return out
