"""Host-side tile-plan search of conv_tc (csrc/conv_tc.cu::conv_tc_plan) over a grid of layer
shapes: every tensor-core-eligible shape must get a plan that fits the opted-in shared memory,
in every math mode and under the plan knobs.  Runs without a GPU (the plan is host code; the
descriptor carries fake, aligned device pointers that are never dereferenced)."""
import ctypes as C
import itertools
import os

import pytest

from atomai_b200 import _C, ops

OPT_IN = 226 * 1024     # cudaFuncAttributeMaxDynamicSharedMemorySize set by ab_conv_tc_fwd


def make_desc(cins, hw, cout, ks, dil, pool, math, affine=True):
    d = _C.Conv()
    d.N, d.H, d.W, d.Cout = 2, hw, hw, cout
    d.ks_h, d.ks_w, d.dil = ks, ks, dil
    d.nsrc = len(cins)
    for i, ci in enumerate(cins):
        e = d.src[i]
        e.ptr = 0x10000000 * (i + 1)
        e.scale = 0x1000 if affine else None
        e.shift = 0x2000 if affine else None
        e.C, e.ld, e.pool = ci, ci, (1 if (pool and i == 0) else 0)
    d.lrelu, d.math, d.out_nchw, d.act = 0.01, math, 0, ops.ACT_LRELU
    return d


def info(d):
    g, b, s = C.c_int(), C.c_int(), C.c_int()
    rc = _C.lib().atomai_b200_conv_info(C.byref(d), C.byref(g), C.byref(b), C.byref(s))
    return rc, g.value, b.value, s.value


SHAPES = [([c], co) for c in (8, 16, 24, 32, 64, 128, 256) for co in (16, 32, 64, 128, 256)] + \
         [([16, 16], 16), ([32, 32], 32), ([64, 64], 64), ([128, 128], 128), ([16, 48], 32)]


@pytest.mark.parametrize("math", [ops.MATH_TF32, ops.MATH_TF32X3])
def test_every_shape_gets_a_plan(math):
    n = 0
    for (cins, cout), ks, dil, pool in itertools.product(SHAPES, (1, 3), (1, 2, 4), (False, True)):
        if ks == 1 and dil > 1:
            continue
        for hw in (8, 64, 512):
            d = make_desc(cins, hw, cout, ks, dil, pool, math)
            rc, grid, block, smem = info(d)
            assert rc == 0, (cins, cout, ks, dil, pool, hw, _C.lib().atomai_b200_last_error())
            assert block == 512 and 1 <= grid <= 4096
            assert 0 < smem <= OPT_IN, (cins, cout, ks, dil, pool, hw, smem)
            n += 1
    assert n >= 900


@pytest.mark.parametrize("env", [{"ATOMAI_B200_NA4": "0", "ATOMAI_B200_SMEM_KB": "212", "ATOMAI_B200_MIN_NR": "2"},
                                 {"ATOMAI_B200_RES_NA4": "1"}, {"ATOMAI_B200_RES_NA4": "0"},
                                 {"ATOMAI_B200_TMA": "0"}, {"ATOMAI_B200_TMA_KC8": "1"}])
def test_plan_knobs_keep_every_unet_layer_plannable(env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    unet = [([16], 32, 3, True), ([32], 32, 3, False), ([32], 64, 3, True), ([64], 64, 3, False),
            ([64], 128, 3, True), ([128], 128, 3, False), ([128], 64, 1, False), ([64, 64], 64, 3, False),
            ([64], 32, 1, False), ([32, 32], 32, 3, False), ([32], 16, 1, False), ([16, 16], 16, 3, False)]
    for cins, cout, ks, pool in unet:
        for hw in (32, 512):
            rc, _, _, smem = info(make_desc(cins, hw, cout, ks, 1, pool, ops.MATH_TF32X3))
            assert rc == 0 and smem <= OPT_IN, (env, cins, cout, smem)
