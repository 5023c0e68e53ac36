"""Launch the fused conv kernel on selected UNet layer shapes (for `ncu --set full`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from atomai_b200 import ops
from atomai_b200.ops import Source
shapes = {"c40": (32, 128, 128, 64), "c1": (32, 512, 1, 16), "c61": (32, 512, 16, 16), "c6": (32, 512, 32, 16), "bn3": (32, 64, 128, 128), "c5": (32, 256, 64, 32), "c51": (32, 256, 32, 32)}
which = [a for a in sys.argv[1:] if a not in ("wgrad", "x3")] or ["c6", "bn3"]
x3 = "x3" in sys.argv[1:]
do_wgrad = "wgrad" in sys.argv[1:]
for tag in which:
    N, hh, cin, cout = shapes[tag]
    dev = "cuda"
    x = torch.rand(N, hh, hh, cin, device=dev); sc = torch.rand(cin, device=dev) + 0.5; sh = torch.rand(cin, device=dev)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05; b = torch.randn(cout, device=dev) * 0.1
    out = torch.empty(N, hh, hh, cout, device=dev); st = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
    math = (ops.MATH_TF32X3 if x3 else ops.MATH_TF32) if cin >= 8 else ops.MATH_FP32
    d = ops.conv_desc([Source(x, sc, sh)], N, hh, hh, cout, (3, 3), 1, 0.01, math)
    wp = ops.prep_weights(w, ops.WMODE_FWD, math)
    for _ in range(3):
        ops.conv_fwd(d, wp, b, out, st)
    if do_wgrad:
        dy = torch.randn(N, hh, hh, cout, device=dev)
        dw = torch.zeros(cout, cin, 3, 3, device=dev)
        for _ in range(3):
            ops.conv_wgrad(d, dy, dw)
    torch.cuda.synchronize()
print("done")
