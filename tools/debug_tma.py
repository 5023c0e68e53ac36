"""TMA-staged vs register-staged activation loaders of conv_tc on the same inputs (bitwise)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from atomai_b200 import ops
from atomai_b200.ops import Source
dev = "cuda"
torch.manual_seed(0)
for (N, H, W, cins, cout, ks, aff) in [(2, 512, 512, [16, 16], 16, 3, True), (2, 512, 512, [16], 32, 3, False),
                                       (2, 256, 256, [32], 32, 3, True), (2, 256, 256, [32, 32], 32, 3, True),
                                       (3, 128, 128, [64], 64, 3, True), (2, 256, 256, [32], 16, 1, True),
                                       (32, 256, 256, [32], 32, 3, True), (2, 40, 24, [32], 32, 3, True)]:
    for math in (ops.MATH_TF32, ops.MATH_TF32X3):
        srcs = []
        for c in cins:
            x = torch.randn(N, H, W, c, device=dev)
            srcs.append(Source(x, torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.3) if aff else Source(x))
        w = torch.randn(cout, sum(cins), ks, ks, device=dev) * 0.1
        b = torch.randn(cout, device=dev) * 0.1
        d = ops.conv_desc(srcs, N, H, W, cout, (ks, ks), 1, 0.01, math)
        wp = ops.prep_weights(w, ops.WMODE_FWD, math)
        outs, stats = [], []
        for tma in ("1", "0"):
            os.environ["ATOMAI_B200_TMA"] = tma
            y = torch.empty(N, H, W, cout, device=dev)
            st = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
            ops.conv_fwd(d, wp, b, y, st)
            torch.cuda.synchronize()
            outs.append(y); stats.append(st)
        diff = (outs[0] - outs[1]).abs()
        bad = diff > 1e-6 * outs[1].abs().max()
        msg = f"{N}x{H}x{W} {cins}->{cout} k{ks} aff={aff} math={math}: maxdiff {float(diff.max()):.2e} bad {int(bad.sum())}"
        if bad.any():
            idx = bad.nonzero()[:5].tolist()
            msg += f" first {idx} n-dist {sorted(set(i[0] for i in bad.nonzero().tolist()))[:8]}"
        print(msg, flush=True)
