"""
Coordinate utilities on the hot path's rim: centre-of-mass atom finder (Locator) and the rVAE
pixel grid / rotation+translation (kept for API parity; the training path fuses the latter into
a CUDA kernel, see atomai_b200/csrc/vae.cu).  atomai/utils/coords.py:21-83.
"""
from typing import Tuple, Union

import numpy as np
import torch
from scipy import ndimage


def find_com(image_data: np.ndarray) -> np.ndarray:
    """Centres of mass of the 4-connected blobs of a thresholded image
    (atomai/utils/coords.py:21-34)."""
    labels, nlabels = ndimage.label(image_data)
    coordinates = np.array(
        ndimage.center_of_mass(image_data, labels, np.arange(nlabels) + 1))
    return coordinates.reshape(coordinates.shape[0], 2)


def grid2xy(X1: torch.Tensor, X2: torch.Tensor) -> torch.Tensor:
    """(M, N) grids -> (M*N, 2) xy coordinates (atomai/utils/coords.py:37-44)."""
    X = torch.stack((X1, X2), 0)
    return X.reshape(2, -1).T


def imcoordgrid(im_dim: Tuple) -> torch.Tensor:
    """Pixel-coordinate grid in [-1, 1]: x (slow axis) ascending, y descending
    (atomai/utils/coords.py:47-54)."""
    xx = torch.linspace(-1, 1, im_dim[0])
    yy = torch.linspace(1, -1, im_dim[1])
    x0, x1 = torch.meshgrid(xx, yy, indexing="ij")
    return grid2xy(x0, x1)


def transform_coordinates(coord: Union[np.ndarray, torch.Tensor],
                          phi: torch.Tensor,
                          coord_dx: Union[np.ndarray, torch.Tensor, int] = 0) -> torch.Tensor:
    """Batched 2-D rotation followed by translation: coord @ [[c, s], [-s, c]] + dx
    (atomai/utils/coords.py:57-83).  Plain torch; used outside the fused training kernel."""
    if isinstance(coord, np.ndarray):
        coord = torch.from_numpy(coord).float()
    if isinstance(coord_dx, np.ndarray):
        coord_dx = torch.from_numpy(coord_dx).float()
    c, s = torch.cos(phi), torch.sin(phi)
    rot = torch.stack([torch.stack([c, s], 1), torch.stack([-s, c], 1)], 1)
    return torch.bmm(coord, rot) + coord_dx


def remove_edge_coord(coordinates: np.ndarray, dim: Tuple, dist_edge: int) -> np.ndarray:
    """Drops the coordinates closer than `dist_edge` to an image edge
    (atomai/utils/coords.py:518-537; note the reference compares column 0 with the WIDTH and
    column 1 with the HEIGHT — kept)."""
    h, w = dim
    c = np.asarray(coordinates)
    if len(c) == 0:
        return coordinates
    bad = (c[:, 0] > w - dist_edge) | (c[:, 0] < dist_edge) | (c[:, 1] > h - dist_edge) | (c[:, 1] < dist_edge)
    return c[~bad]
