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


# ------------------------------------------------------------------ sub-image extraction
# the step between Segmentor.predict and the VAE path
# (atomai/utils/img.py:138-180, 183-350, 502-551).  The reference crops window by window in a
# Python loop; here every window of an image is one vectorised gather.  Semantics that tests
# pin against the reference: np.around (half-to-even) centre rounding, windows that do not fit
# are dropped, so are windows containing NaN, numpy's negative-index wrap-around is honoured
# (a window whose start AND stop are negative is a legal crop from the far edge), coordinates
# are returned un-rounded, empty results are (None, None) / three empty lists.
def _window_starts(centres: np.ndarray, r: int, dim: int):
    """start index of each window along one axis (after numpy's negative-index rules) and a mask
    of the windows whose slice has exactly r elements."""
    half = r // 2
    start = centres - half
    stop = centres + half + (1 if r % 2 else 0)
    ok = (start >= 0) & (stop <= dim)
    wrap = (stop < 0) & (start + dim >= 0)          # both negative: slice counted from the end
    start = np.where(wrap, start + dim, start)
    return start, ok | wrap


def get_imgstack(imgdata: np.ndarray, coord: np.ndarray, r: int):
    """Sub-images of side `r` centred at the (x, y) coordinates of ONE image (h, w[, c]);
    returns (stack, kept coordinates) or (None, None) — atomai/utils/img.py:138-180."""
    coord = np.asarray(coord)
    if len(coord) == 0:
        return None, None
    r = int(r)
    cx = np.around(coord[:, 0]).astype(np.int64)
    cy = np.around(coord[:, 1]).astype(np.int64)
    sx, okx = _window_starts(cx, r, imgdata.shape[0])
    sy, oky = _window_starts(cy, r, imgdata.shape[1])
    keep = okx & oky
    if not keep.any():
        return None, None
    idx = np.nonzero(keep)[0]
    ar = np.arange(r)
    rows = (sx[idx, None] + ar[None, :])[:, :, None]            # (k, r, 1)
    cols = (sy[idx, None] + ar[None, :])[:, None, :]            # (k, 1, r)
    stack = imgdata[rows, cols]                                 # (k, r, r[, c]) — one gather
    if np.issubdtype(stack.dtype, np.floating):
        finite = ~np.isnan(stack.reshape(len(idx), -1)).any(1)
        if not finite.any():
            return None, None
        stack, idx = stack[finite], idx[finite]
    return np.ascontiguousarray(stack), coord[idx]


def extract_subimages_cuda(imgdata, coordinates, window_size: int, coord_class: int = 0):
    """extract_subimages for a CUDA image stack (n, h, w, c) fp32: the window bookkeeping (centre
    rounding, which windows fit, numpy's negative-index wrap) is the same host code as the numpy
    path, the crops themselves are ONE gather kernel over all frames
    (atomai_b200_gather_windows) that also flags windows containing NaN.  Returns
    (CUDA tensor (K, r, r, c), centres (K, 2) numpy, frame index (K,) numpy), or three empty
    lists when nothing fits (like the reference)."""
    import torch
    from .. import ops
    assert imgdata.is_cuda and imgdata.dim() == 4 and imgdata.dtype == torch.float32
    if isinstance(coordinates, np.ndarray):
        coordinates = {0: np.concatenate((coordinates, np.zeros((coordinates.shape[0], 1))), axis=-1)}
    n, h, w, c = imgdata.shape
    r = int(window_size)
    rows, coms, frames = [], [], []
    for i, coord in zip(range(n), coordinates.values()):
        coord_i = np.asarray(coord)
        coord_i = coord_i[coord_i[:, 2] == coord_class][:, :2]
        if len(coord_i) == 0:
            continue
        cx = np.around(coord_i[:, 0]).astype(np.int64)
        cy = np.around(coord_i[:, 1]).astype(np.int64)
        sx, okx = _window_starts(cx, r, h)
        sy, oky = _window_starts(cy, r, w)
        idx = np.nonzero(okx & oky)[0]
        if len(idx) == 0:
            continue
        rows.append(np.stack([np.full(len(idx), i), sx[idx], sy[idx]], 1))
        coms.append(coord_i[idx])
        frames.append(np.ones(len(idx), int) * i)
    if not rows:
        return [], [], []
    table = torch.from_numpy(np.concatenate(rows).astype(np.int32)).to(imgdata.device)
    K = table.shape[0]
    out = torch.empty((K, r, r, c), device=imgdata.device, dtype=torch.float32)
    nanflag = torch.zeros(K, device=imgdata.device, dtype=torch.int32)
    ops.gather_windows(imgdata.contiguous(), table, r, out, nanflag)
    keep = (nanflag == 0).cpu().numpy()
    coms, frames = np.concatenate(coms), np.concatenate(frames)
    if not keep.all():
        if not keep.any():
            return [], [], []
        out = out[torch.from_numpy(keep).to(out.device)]
        coms, frames = coms[keep], frames[keep]
    return out, coms, frames


def extract_subimages(imgdata: np.ndarray, coordinates, window_size: int, coord_class: int = 0):
    """(sub-images, centres, frame index) around the atoms of class `coord_class` in a stack
    (n, h, w, c) with Locator-style coordinates {i: (N, 3)} — atomai/utils/img.py:298-350.
    Images are paired with the dictionary's VALUES in insertion order, like the reference.
    A CUDA tensor stack takes the gather-kernel path (extract_subimages_cuda)."""
    if hasattr(imgdata, "is_cuda") and imgdata.is_cuda:
        if imgdata.dim() == 2:
            imgdata = imgdata[None, ..., None]
        return extract_subimages_cuda(imgdata, coordinates, window_size, coord_class)
    if isinstance(coordinates, np.ndarray):
        coordinates = {0: np.concatenate((coordinates, np.zeros((coordinates.shape[0], 1))), axis=-1)}
    if np.ndim(imgdata) == 2:
        imgdata = imgdata[None, ..., None]
    subimages_all, com_all, frames_all = [], [], []
    for i, (img, coord) in enumerate(zip(imgdata, coordinates.values())):
        coord_i = coord[coord[:, 2] == coord_class][:, :2]
        stack_i, com_i = get_imgstack(img, coord_i, window_size)
        if stack_i is None:
            continue
        subimages_all.append(stack_i)
        com_all.append(com_i)
        frames_all.append(np.ones(len(com_i), int) * i)
    if len(subimages_all) > 0:
        subimages_all = np.concatenate(subimages_all, axis=0)
        com_all = np.concatenate(com_all, axis=0)
        frames_all = np.concatenate(frames_all, axis=0)
    return subimages_all, com_all, frames_all


def imcrop_randpx(img: np.ndarray, window_size: int, num_images: int, random_state: int = 0):
    """`num_images` windows at distinct random pixels (np.random global state, same draw order as
    atomai/utils/img.py:183-211; `random_state` is accepted and ignored there too)."""
    seen, com = set(), []
    lo = window_size // 2 + 1
    while len(com) < num_images:
        x = np.random.randint(lo, img.shape[0] - window_size // 2 - 1)
        y = np.random.randint(lo, img.shape[1] - window_size // 2 - 1)
        if (x, y) not in seen:
            seen.add((x, y))
            com.append((x, y))
    return get_imgstack(img, np.array(com), window_size)


def imcrop_randcoord(img: np.ndarray, coord: np.ndarray, window_size: int, num_images: int,
                     random_state: int = 0):
    """`num_images` windows at distinct randomly chosen coordinates (atomai/utils/img.py:214-234)."""
    seen, com = set(), []
    while len(com) < num_images:
        i = np.random.randint(len(coord))
        if i not in seen:
            seen.add(i)
            com.append(coord[i].tolist())
    return get_imgstack(img, np.array(com), window_size)


def extract_random_subimages(imgdata: np.ndarray, window_size: int, num_images: int,
                             coordinates=None, **kwargs: int):
    """`num_images` random windows per frame, at atoms of `coord_class` or at random pixels
    (atomai/utils/img.py:237-295)."""
    from .coords import remove_edge_coord
    coord_class = kwargs.get("coord_class", 0)
    if np.ndim(imgdata) < 4:
        imgdata = imgdata[..., None]
    n = imgdata.shape[0]
    subimages_all = np.zeros((num_images * n, window_size, window_size, imgdata.shape[-1]))
    com_all = np.zeros((num_images * n, 2))
    frames_all = np.zeros((num_images * n))
    for i, img in enumerate(imgdata):
        if not coordinates:
            stack_i, com_i = imcrop_randpx(img, window_size, num_images, random_state=i)
        else:
            coord = coordinates[i]
            coord = coord[coord[:, -1] == coord_class][:, :2]
            coord = remove_edge_coord(coord, imgdata.shape[1:3], window_size // 2 + 1)
            if num_images > len(coord):
                raise ValueError("Number of images cannot be greater than the available coordinates")
            stack_i, com_i = imcrop_randcoord(img, coord, window_size, num_images, random_state=i)
        sl = slice(i * num_images, (i + 1) * num_images)
        subimages_all[sl] = stack_i
        com_all[sl] = com_i
        frames_all[sl] = np.ones(len(com_i), int) * i
    return subimages_all, com_all, frames_all


def crop_borders(imgdata: np.ndarray, thresh: float = 0) -> np.ndarray:
    """Drops the rows/columns of (h, w, c) whose values are all <= thresh, channel by channel
    (atomai/utils/img.py:502-519)."""
    out = []
    for i in range(imgdata.shape[-1]):
        img = imgdata[..., i]
        mask = img > thresh
        out.append(img[np.ix_(mask.any(1), mask.any(0))])
    return np.array(out).transpose(1, 2, 0)


def get_coord_grid(imgdata: np.ndarray, step: int, return_dict: bool = True):
    """Square grid of (row, col) points with spacing `step` for every image of a stack, in the
    Locator dictionary format or as one array (atomai/utils/img.py:522-551)."""
    if np.ndim(imgdata) == 2:
        imgdata = np.expand_dims(imgdata, axis=0)
    ii, jj = np.meshgrid(np.arange(0, imgdata.shape[1], step), np.arange(0, imgdata.shape[2], step),
                         indexing="ij")
    coord = np.stack([ii.ravel(), jj.ravel()], axis=1)
    if return_dict:
        coord = np.concatenate((coord, np.zeros((coord.shape[0], 1))), axis=-1)
        return {i: coord for i in range(imgdata.shape[0])}
    return np.concatenate([coord for _ in range(imgdata.shape[0])], axis=0)
