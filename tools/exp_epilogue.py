"""Where does conv_tc's time go?  Thin layers with the epilogue's / loader's parts switched off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from atomai_b200 import ops
from atomai_b200.ops import Source
dev = "cuda"
def timeit(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (N, hh, cin, cout) in [(32, 256, 32, 32), (32, 512, 16, 16), (32, 128, 64, 64)]:
    x = torch.rand(N, hh, hh, cin, device=dev)
    sc = torch.rand(cin, device=dev) + 0.5; sh = torch.rand(cin, device=dev)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05; b = torch.randn(cout, device=dev) * 0.1
    out = torch.empty(N, hh, hh, cout, device=dev)
    st = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
    for math, name in ((ops.MATH_TF32, "tf32"), (ops.MATH_TF32X3, "x3")):
        wp = ops.prep_weights(w, ops.WMODE_FWD, math)
        row = f"{N}x{hh}^2 {cin}->{cout} {name}:"
        for tag, kw in (("full", dict(aff=True, stats=True)), ("no-stats", dict(aff=True, stats=False)),
                        ("no-affine", dict(aff=False, stats=True)), ("neither", dict(aff=False, stats=False))):
            src = Source(x, sc, sh) if kw["aff"] else Source(x)
            d = ops.conv_desc([src], N, hh, hh, cout, (3, 3), 1, 0.01, math)
            t = timeit(lambda: ops.conv_fwd(d, wp, b, out, st if kw["stats"] else None))
            row += f"  {tag} {t:6.1f}"
        for tma in ("1", "0"):
            os.environ["ATOMAI_B200_TMA"] = tma
            d = ops.conv_desc([Source(x, sc, sh)], N, hh, hh, cout, (3, 3), 1, 0.01, math)
            row += f"  tma{tma} {timeit(lambda: ops.conv_fwd(d, wp, b, out, st)):6.1f}"
        os.environ.pop("ATOMAI_B200_TMA", None)
        for dbg in ("1", "2", "3"):
            os.environ["ATOMAI_B200_DBG"] = dbg
            d = ops.conv_desc([Source(x, sc, sh)], N, hh, hh, cout, (3, 3), 1, 0.01, math)
            row += f"  dbg{dbg} {timeit(lambda: ops.conv_fwd(d, wp, b, out, None)):6.1f}"
        os.environ.pop("ATOMAI_B200_DBG", None)
        print(row, flush=True)
