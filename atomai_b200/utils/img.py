"""
Image helpers on the prediction path (host-side numpy): padding to the network's downsampling
factor, binary thresholding and resizing — atomai/utils/img.py:20-40, 112-135, 554-564.
"""
from typing import Tuple

import numpy as np


def img_pad(image_data: np.ndarray, pooling: int) -> np.ndarray:
    """Zero-pads (n, h, w) at the bottom/right so that h and w are divisible by `pooling`
    (atomai/utils/img.py:112-135; float64 result like the reference's np.concatenate)."""
    pooling = int(pooling)
    n, h, w = image_data.shape
    ph, pw = (-h) % pooling, (-w) % pooling
    if ph == 0 and pw == 0:
        return image_data
    out = np.zeros((n, h + ph, w + pw), dtype=np.result_type(image_data.dtype, np.float64))
    out[:, :h, :w] = image_data
    return out


def cv_thresh(imgdata: np.ndarray, threshold: float = .5) -> np.ndarray:
    """cv2.threshold(img, threshold, 1, THRESH_BINARY): 1.0 where img > threshold else 0.0, same
    dtype as the input (atomai/utils/img.py:554-564).  Pure numpy so cv2 is not a dependency."""
    return (imgdata > threshold).astype(imgdata.dtype)


def img_resize(image_data: np.ndarray, rs: Tuple[int], round_: bool = False) -> np.ndarray:
    """Resizes a stack of images (atomai/utils/img.py:20-40); needs cv2 like the reference."""
    import cv2
    if rs[0] != rs[1]:
        rs = (rs[1], rs[0])
    if image_data.shape[1:3] == rs:
        return image_data.copy()
    out = np.zeros((image_data.shape[0], rs[0], rs[1]))
    inter = cv2.INTER_AREA if image_data.shape[1] > rs[0] else cv2.INTER_CUBIC
    for i, img in enumerate(image_data):
        img = cv2.resize(img, (rs[0], rs[1]), interpolation=inter)
        out[i] = np.round(img) if round_ else img
    return out
