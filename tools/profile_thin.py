"""First layer (1 -> 16, 3x3) and head (16 -> 3, 1x1) forward / gradients at the bench shape, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from atomai_b200 import ops
from atomai_b200.ops import Source
dev = "cuda"
N, hh = 32, 512
x = torch.rand(N, hh, hh, 1, device=dev)
w = torch.randn(16, 1, 3, 3, device=dev) * 0.1; b = torch.randn(16, device=dev) * 0.1
out = torch.empty(N, hh, hh, 16, device=dev); st = torch.zeros(32, device=dev, dtype=torch.float64)
d = ops.conv_desc([Source(x)], N, hh, hh, 16, (3, 3), 1, 0.01, ops.MATH_FP32)
wp = ops.prep_weights(w, ops.WMODE_FWD, ops.MATH_FP32)
for _ in range(2):
    ops.conv_fwd(d, wp, b, out, st)
dy = torch.randn(N, hh, hh, 16, device=dev); dw = torch.zeros(16, 1, 3, 3, device=dev)
for _ in range(2):
    ops.conv_wgrad(d, dy, dw)
# head: 16 -> 3 with the BN affine of c6 pending
sc = torch.rand(16, device=dev) + 0.5; sh = torch.rand(16, device=dev)
wpx = torch.randn(3, 16, 1, 1, device=dev) * 0.1; bpx = torch.randn(3, device=dev) * 0.1
dpx = ops.conv_desc([Source(out, sc, sh)], N, hh, hh, 3, (1, 1), 1, 1.0, ops.MATH_FP32)
wpp = ops.prep_weights(wpx, ops.WMODE_FWD, ops.MATH_FP32)
logits = torch.empty(N, hh, hh, 3, device=dev)
for _ in range(2):
    ops.conv_fwd(dpx, wpp, bpx, logits, None)
# its data gradient: 3 -> 16
dl = torch.randn(N, hh, hh, 3, device=dev)
dd = ops.conv_desc([Source(dl)], N, hh, hh, 16, (1, 1), 1, 1.0, ops.MATH_FP32)
wd = ops.prep_weights(wpx, ops.WMODE_DGRAD, ops.MATH_FP32)
dx = torch.empty(N, hh, hh, 16, device=dev)
for _ in range(2):
    ops.conv_fwd(dd, wd, None, dx, None)
dwp = torch.zeros(3, 16, 1, 1, device=dev)
for _ in range(2):
    ops.conv_wgrad(dpx, dl, dwp)
torch.cuda.synchronize()
print("done")
