"""Sub-image extraction (the step between Segmentor.predict and the VAE path)
against goldens produced by the unmodified reference (tests/golden/make_golden_img.py):
atomai/utils/img.py:138-350, 502-551; atomai/utils/coords.py:518-537.  Index work: bit-exact."""
import numpy as np

import golden_utils as gu
from atomai_b200.utils import (crop_borders, extract_random_subimages, extract_subimages,
                               get_coord_grid, get_imgstack, remove_edge_coord)


def _eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    assert np.array_equal(a, b, equal_nan=True)


def test_extract_subimages_matches_reference():
    g = gu.load("subimages.npz")
    stack, coords, single, single_xy = gu.subimage_inputs()
    for r in (7, 8, 1, 12):
        for cls, tag in ((0, f"dict_r{r}"), (1, f"dict_c1_r{r}")):
            s, c, f = extract_subimages(stack, coords, r, coord_class=cls)
            _eq(s, g[tag + "/sub"]); _eq(c, g[tag + "/com"]); _eq(f, g[tag + "/frames"])
    s, c, f = extract_subimages(single, single_xy, 9)
    _eq(s, g["single_r9/sub"]); _eq(c, g["single_r9/com"]); _eq(f, g["single_r9/frames"])
    s, c = get_imgstack(stack[0], coords[0][:, :2], 6)
    _eq(s, g["stack_r6/sub"]); _eq(c, g["stack_r6/com"])


def test_nothing_fits_and_empty_inputs():
    stack, coords, _, _ = gu.subimage_inputs()
    assert get_imgstack(stack[0], coords[0][:, :2], 64) == (None, None)
    assert get_imgstack(stack[0], np.zeros((0, 2)), 5) == (None, None)
    s, c, f = extract_subimages(stack, coords, 64)
    assert s == [] and c == [] and f == []          # the reference returns three empty lists


def test_random_subimages_same_draw_order():
    g = gu.load("subimages.npz")
    stack, coords, _, _ = gu.subimage_inputs()
    clean = np.nan_to_num(stack)
    np.random.seed(7)
    s, c, f = extract_random_subimages(clean, 8, 5)
    _eq(s, g["rand_px/sub"]); _eq(c, g["rand_px/com"]); _eq(f, g["rand_px/frames"])
    np.random.seed(11)
    s, c, f = extract_random_subimages(clean[:, :36, :36], 6, 3, coordinates=coords, coord_class=0)
    _eq(s, g["rand_coord/sub"]); _eq(c, g["rand_coord/com"]); _eq(f, g["rand_coord/frames"])


def test_grid_edges_and_border_crop():
    g = gu.load("subimages.npz")
    stack, coords, single, _ = gu.subimage_inputs()
    _eq(remove_edge_coord(coords[1][:, :2], stack.shape[1:3], 5), g["edge/kept"])
    _eq(get_coord_grid(single, 7)[0], g["grid/dict0"])
    _eq(get_coord_grid(stack[..., 0], 9, return_dict=False), g["grid/arr"])
    enc = -1e5 * np.ones((20, 18, 2))
    enc[3:15, 4:16] = gu.images(5, 1, 12, 12)[0][..., None] + 1.0
    _eq(crop_borders(enc, -1e5), g["crop/out"])


def test_large_stack_gather_property():
    """Size-independent property at a realistic size: every returned window equals the direct slice
    of its (rounded) centre, and all in-bounds, NaN-free centres are returned in order."""
    rs = np.random.RandomState(0)
    img = rs.rand(1024, 1024, 3).astype(np.float32)
    xy = rs.rand(5000, 2) * 1024
    r = 32
    sub, com = get_imgstack(img, xy, r)
    cx, cy = np.around(xy[:, 0]).astype(int), np.around(xy[:, 1]).astype(int)
    ok = (cx - 16 >= 0) & (cx + 16 <= 1024) & (cy - 16 >= 0) & (cy + 16 <= 1024)
    assert np.array_equal(com, xy[ok]) and sub.shape == (int(ok.sum()), r, r, 3)
    for k in rs.choice(len(com), 50, replace=False):
        a, b = int(np.around(com[k, 0])), int(np.around(com[k, 1]))
        assert np.array_equal(sub[k], img[a - 16:a + 16, b - 16:b + 16])
