"""Print the conv_tc tile plan (`ATOMAI_B200_PLAN_LOG`) of every tensor-core layer of the default
Unet, forward and dgrad, without a GPU: the plan search is host code and the descriptor's device
pointers are never dereferenced.  usage: [M=tf32|tf32x3] [ATOMAI_B200_NA4=..] show_plans.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atomai_b200 import _C, ops

LAYERS = [  # name, H, cins, cout, ks, pooled first source
    ("c2.0", 256, [16], 32, 3, True), ("c2.1", 256, [32], 32, 3, False),
    ("c3.0", 128, [32], 64, 3, True), ("c3.1", 128, [64], 64, 3, False),
    ("bn.0", 64, [64], 128, 3, True), ("bn.1", 64, [128], 128, 3, False),
    ("u1", 64, [128], 64, 1, False), ("c4.0", 128, [64, 64], 64, 3, False), ("c4.1", 128, [64], 64, 3, False),
    ("u2", 128, [64], 32, 1, False), ("c5.0", 256, [32, 32], 32, 3, False), ("c5.1", 256, [32], 32, 3, False),
    ("u3", 256, [32], 16, 1, False), ("c6.0", 512, [16, 16], 16, 3, False),
]
MATH = {"tf32": ops.MATH_TF32, "tf32x3": ops.MATH_TF32X3}[os.environ.get("M", "tf32x3")]
os.environ["ATOMAI_B200_PLAN_LOG"] = "1"
lib = _C.lib()


def desc(cins, hh, cout, ks, pool, aff):
    d = _C.Conv()
    d.N, d.H, d.W, d.Cout = 32, hh, hh, cout
    d.ks_h, d.ks_w, d.dil = ks, ks, 1
    d.nsrc = len(cins)
    for i, ci in enumerate(cins):
        e = d.src[i]
        e.ptr = 0x10000000 * (i + 1)
        e.scale = 0x1000 if aff else None
        e.shift = 0x2000 if aff else None
        e.C, e.ld, e.pool = ci, ci, (1 if pool else 0)
    d.lrelu, d.math, d.out_nchw, d.act = 0.01, MATH, 0, ops.ACT_LRELU
    return d


def show(tag, d):
    sys.stderr.write(f"{tag:12s}")
    sys.stderr.flush()
    g, b, s = C.c_int(), C.c_int(), C.c_int()
    if lib.atomai_b200_conv_info(C.byref(d), C.byref(g), C.byref(b), C.byref(s)):
        sys.stderr.write("ERR " + lib.atomai_b200_last_error().decode() + "\n")


for name, hh, cins, cout, ks, pool in LAYERS:
    show(f"{name} fwd", desc(cins, hh, cout, ks, pool, True))
    if not pool:
        show(f"{name} dgrad", desc([cout], hh, sum(cins), ks, False, False))
