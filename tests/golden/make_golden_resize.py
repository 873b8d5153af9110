"""cv2.INTER_AREA fixture for ofb_area_resize_u8 (authoring container only; needs cv2):
the RGB path of the reference's loader, dataset_loader_stanford.py:92-94."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(11)
img = rng.integers(0, 256, size=(64, 96, 3), dtype=np.uint8)
img[:8, :8] = 255            # saturated block: sums at the top of the range
img[8:16, :8] = np.arange(8 * 8 * 3, dtype=np.uint8).reshape(8, 8, 3) % 7   # many .5 ties after /4 and /16
out = {"src": img}
for f in (2, 4, 8):
    out[f"area_{f}"] = cv2.resize(img, (96 // f, 64 // f), interpolation=cv2.INTER_AREA)
np.savez_compressed(os.path.join(HERE, "area_resize.npz"), **out)
print({k: v.shape for k, v in out.items()}, cv2.__version__)
