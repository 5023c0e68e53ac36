"""Atom-coordinate extraction is bit-exact with the reference: Locator (product, scipy-based like
the reference) and the independent numpy oracle both reproduce coordinates the unmodified reference
produced on a crop of its own golden NN output (test/predictors/test_locator.py:20-39)."""
import numpy as np

import golden_utils as gu
from atomai_b200.predictors import Locator
from oracle.locator_ref import locate


def test_locator_bit_exact_channel_last():
    g = gu.load("locator_crop.npz")
    got = Locator(0.5, 5).run(g["nn_output"])
    assert got[0].dtype == np.float64
    assert np.array_equal(got[0], g["coordinates"])


def test_locator_bit_exact_channel_first():
    g = gu.load("locator_crop.npz")
    got = Locator(0.5, 5, dim_order="channel_first").run(np.transpose(g["nn_output"], (0, 3, 1, 2)))
    assert np.array_equal(got[0], g["coordinates"])


def test_oracle_locator_matches_reference_golden():
    g = gu.load("locator_crop.npz")
    assert np.array_equal(locate(g["nn_output"])[0], g["coordinates"])


def test_locator_single_channel_and_empty():
    out = np.zeros((1, 32, 32, 1), np.float32)
    assert Locator().run(out)[0].shape == (0, 3)
    out[0, 10:13, 10:13, 0] = 0.9
    c = Locator().run(out)[0]
    assert c.shape == (1, 3) and np.array_equal(c[0], [11.0, 11.0, 0.0])
    out[0, 1:3, 1:3, 0] = 0.9          # inside the 5 px edge band -> dropped
    assert Locator().run(out)[0].shape == (1, 3)
