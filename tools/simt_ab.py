"""Old vs new thin-channel SIMT kernels on the same inputs (env switches ATOMAI_B200_C1_TILE /
ATOMAI_B200_PIX_QUAD are read at every launch): first layer 1 -> 16 forward + weight gradient
(tile kernels), head data gradient nb_classes -> 16 and head weight gradient 16 -> nb_classes
(quad-lane kernels).  Prints max |new - old| (forward kernels keep the per-output fmaf order:
expected 0) and CUDA-event times at the bench shape (batch 32 x 512^2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from atomai_b200 import ops
from atomai_b200.ops import Source
dev = "cuda"


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def both(key, run, time_it):
    res = {}
    for tag, v in (("old", "0"), ("new", "1")):
        os.environ[key] = v
        out = run()
        torch.cuda.synchronize()
        res[tag] = ([o.clone() for o in out], timeit(run) if time_it else 0.0)
    os.environ.pop(key, None)
    return res


def report(name, res, rtol):
    worst = 0.0
    for a, b in zip(res["old"][0], res["new"][0]):
        den = max(a.abs().max().item(), 1e-30)
        worst = max(worst, (a.double() - b.double()).abs().max().item() / den)
    flag = "ok" if worst <= rtol else "MISMATCH"
    print(f"{name:58s} rel diff {worst:9.2e} {flag:8s} old {res['old'][1]:7.1f} us  new {res['new'][1]:7.1f} us", flush=True)
    return worst <= rtol


ok = True
torch.manual_seed(0)
for (N, H, W, cout, aff, big) in [(32, 512, 512, 16, False, True), (3, 37, 120, 16, True, False),
                                  (2, 64, 64, 32, True, False), (1, 16, 64, 16, False, False),
                                  (2, 33, 250, 16, True, False)]:
    x = torch.rand(N, H, W, 1, device=dev)
    sc = (torch.rand(1, device=dev) + 0.5) if aff else None
    sh = (torch.rand(1, device=dev) - 0.5) if aff else None
    w = torch.randn(cout, 1, 3, 3, device=dev) * 0.3
    b = torch.randn(cout, device=dev) * 0.1
    out = torch.empty(N, H, W, cout, device=dev)
    st = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
    d = ops.conv_desc([Source(x, sc, sh)], N, H, W, cout, (3, 3), 1, 0.01, ops.MATH_FP32)
    wp = ops.prep_weights(w, ops.WMODE_FWD, ops.MATH_FP32)

    def fwd():
        st.zero_()
        ops.conv_fwd(d, wp, b, out, st)
        return [out, st]
    r = both("ATOMAI_B200_C1_TILE", fwd, big)
    ok &= report(f"conv_c1 fwd  N{N} {H}x{W} 1->{cout} aff={aff}", r, 1e-6)
    if cout == 16:
        dy = torch.randn(N, H, W, cout, device=dev)
        dw = torch.zeros(cout, 1, 3, 3, device=dev)

        def wg():
            dw.zero_()
            ops.conv_wgrad(d, dy, dw)
            return [dw]
        r = both("ATOMAI_B200_C1_TILE", wg, big)
        ok &= report(f"conv_c1 wgrad N{N} {H}x{W} 1->{cout} aff={aff}", r, 2e-4)
    del x, out

for (N, H, W, ncls, aff, big) in [(32, 512, 512, 3, True, True), (3, 37, 53, 3, False, False),
                                  (2, 40, 40, 1, True, False), (2, 31, 17, 4, True, False)]:
    # head data gradient: ncls -> 16 (1x1), identity activation
    dl = torch.randn(N, H, W, ncls, device=dev)
    wpx = torch.randn(ncls, 16, 1, 1, device=dev) * 0.3
    dd = ops.conv_desc([Source(dl)], N, H, W, 16, (1, 1), 1, 1.0, ops.MATH_FP32)
    wd = ops.prep_weights(wpx, ops.WMODE_DGRAD, ops.MATH_FP32)
    dx = torch.empty(N, H, W, 16, device=dev)

    def dg():
        ops.conv_fwd(dd, wd, None, dx, None)
        return [dx]
    r = both("ATOMAI_B200_PIX_QUAD", dg, big)
    ok &= report(f"head dgrad   N{N} {H}x{W} {ncls}->16", r, 1e-6)
    # head weight gradient: 16 -> ncls with the pending BN affine of the last block
    a = torch.rand(N, H, W, 16, device=dev)
    sc = (torch.rand(16, device=dev) + 0.5) if aff else None
    sh = (torch.rand(16, device=dev) - 0.5) if aff else None
    dpx = ops.conv_desc([Source(a, sc, sh)], N, H, W, ncls, (1, 1), 1, 1.0, ops.MATH_FP32)
    dwp = torch.zeros(ncls, 16, 1, 1, device=dev)

    def hw():
        dwp.zero_()
        ops.conv_wgrad(dpx, dl, dwp)
        return [dwp]
    r = both("ATOMAI_B200_PIX_QUAD", hw, big)
    ok &= report(f"head wgrad   N{N} {H}x{W} 16->{ncls} aff={aff}", r, 2e-4)
    del dl, dx, a
print("ALL OK" if ok else "SOME MISMATCH")
sys.exit(0 if ok else 1)
