"""Device side of the reference's Dataset.__getitem__ (dataset_loader_stanford.py:52,79): the decoded uint8 panorama
(cv2 channel order, HWC) becomes the float32 CHW network input, `rgb.astype(np.float32) / 255` then
`transpose(2, 0, 1)`, and the cv2.INTER_AREA down-scaling in front of it (:92-96) for integer factors.  Decoding
stays on the host (cv2.imread); shipping the uint8 image and converting on the GPU moves 4x fewer bytes over PCIe
than the reference's float32 batch."""
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


def area_resize_u8(img_u8, factor):
    """cv2.resize(img, (W // factor, H // factor), interpolation=cv2.INTER_AREA) on the device for an integer
    factor (dataset_loader_stanford.py:92-96: 2048x4096 Stanford panoramas to 1024x2048 / 512x1024).
    img_u8 (B,H,W,C) or (H,W,C) uint8 CUDA -> (B,H/f,W/f,C) uint8, bit-identical to OpenCV."""
    if img_u8.dim() == 3:
        img_u8 = img_u8.unsqueeze(0)
    if not img_u8.is_cuda or img_u8.dtype != torch.uint8 or img_u8.dim() != 4:
        raise _lib.OfbError(f"img_u8 must be a uint8 (B,H,W,C) CUDA tensor, got {img_u8.dtype} {tuple(img_u8.shape)} on {img_u8.device}")
    img_u8 = img_u8.contiguous()
    b, h, w, c = img_u8.shape
    out = torch.empty(b, h // factor, w // factor, c, dtype=torch.uint8, device=img_u8.device)
    _lib.use_device(img_u8.device)
    _lib.check(_lib.lib().ofb_area_resize_u8(_lib.ptr(img_u8), b, h, w, c, int(factor), _lib.ptr(out),
                                             _lib.stream_of(img_u8.device)))
    return out


def load_panorama_batch(imgs_u8, factor=1):
    """The loader's RGB path on the device (dataset_loader_stanford.py:92-94 + :52,79): INTER_AREA resize by an
    integer factor, then float32 CHW / 255."""
    x = area_resize_u8(imgs_u8, factor) if factor != 1 else imgs_u8
    return rgb_u8_to_input(x)
