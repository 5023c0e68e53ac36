"""Plan sweep for the streamed-weight layers: time conv fwd under ATOMAI_B200_PLAN overrides."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, torch
sys.path.insert(0, %r)
from atomai_b200 import ops
from atomai_b200.ops import Source
N = 32
LAY = [("bn.0", 64, [64], 128, True), ("bn.1", 64, [128], 128, False), ("c4.0", 128, [64, 64], 64, False), ("c3.1", 128, [64], 64, False), ("c5.0", 256, [32, 32], 32, False), ("dg_bn", 64, [128], 128, False)]
for name, hh, cins, cout, pool in LAY:
    srcs = []
    for ci in cins:
        s = 2 * hh if pool else hh
        srcs.append(Source(torch.rand(N, s, s, ci, device="cuda"), torch.rand(ci, device="cuda") + .5, torch.rand(ci, device="cuda"), pool))
    cin = sum(cins)
    w = torch.randn(cout, cin, 3, 3, device="cuda") * .05; b = torch.randn(cout, device="cuda")
    out = torch.empty(N, hh, hh, cout, device="cuda"); st = torch.zeros(2 * cout, device="cuda", dtype=torch.float64)
    d = ops.conv_desc(srcs, N, hh, hh, cout, (3, 3), 1, 0.01, ops.MATH_TF32)
    wp = ops.prep_weights(w, ops.WMODE_FWD, ops.MATH_TF32)
    try:
        for _ in range(2): ops.conv_fwd(d, wp, b, out, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.conv_fwd(d, wp, b, out, st)
        e1.record(); torch.cuda.synchronize()
        print(f"{name}:{e0.elapsed_time(e1)/5*1e3:.0f}", end=" ")
    except Exception as ex:
        print(f"{name}:--", end=" ")
print()
''' % ROOT
for plan in ["", "1,32", "1,16", "1,8", "2,32", "2,16", "2,8", "0,32", "0,16"]:
    env = dict(os.environ)
    if plan: env["ATOMAI_B200_PLAN"] = plan
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(f"plan {plan or 'auto':6s}", r.stdout.strip() or r.stderr.strip()[-200:], flush=True)
