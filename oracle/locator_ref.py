"""
TEST INFRASTRUCTURE — independent numpy restatement of the reference's Locator
(atomai/predictors/predictor.py:582-639 + utils/img.py:554-564 + utils/coords.py:21-34):
threshold (x > t) -> 4-connected components (explicit flood fill, scipy.ndimage.label's default
structure and raster label order) -> centre of mass of the 0/1 image (sums of integer coordinates
are exact in float64, so any summation order gives the same bits) -> edge filter.
Pinned by tests/golden/locator_crop.npz (coordinates produced by the unmodified reference).
"""
import numpy as np


def label4(mask: np.ndarray):
    h, w = mask.shape
    lab = np.zeros((h, w), np.int32)
    n = 0
    for i in range(h):
        for j in range(w):
            if mask[i, j] and lab[i, j] == 0:
                n += 1
                stack = [(i, j)]
                lab[i, j] = n
                while stack:
                    a, b = stack.pop()
                    for da, db in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                        x, y = a + da, b + db
                        if 0 <= x < h and 0 <= y < w and mask[x, y] and lab[x, y] == 0:
                            lab[x, y] = n
                            stack.append((x, y))
    return lab, n


def locate(nn_output: np.ndarray, threshold: float = 0.5, dist_edge: int = 5):
    """nn_output (n, h, w, c) channel-last -> {i: (n_atoms, 3) float64}."""
    if nn_output.shape[-1] == 1:
        nn_output = np.concatenate((nn_output, 1 - nn_output), axis=3)
    n, h, w, c = nn_output.shape
    out = {}
    for i in range(n):
        rows = []
        for ch in range(c - 1):
            mask = nn_output[i, :, :, ch] > threshold
            lab, k = label4(mask)
            for l in range(1, k + 1):
                ys, xs = np.nonzero(lab == l)
                cy = np.float64(ys.sum()) / np.float64(len(ys))
                cx = np.float64(xs.sum()) / np.float64(len(xs))
                if cy > h - dist_edge or cy < dist_edge or cx > w - dist_edge or cx < dist_edge:
                    continue
                rows.append((cy, cx, float(ch)))
        out[i] = np.array(rows, np.float64).reshape(-1, 3)
    return out
