"""Debug: per-layer comparison of the tf32 vs fp32 math modes in train-mode forward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import atomai_b200 as ab
from atomai_b200 import engine
from test_oracle import build_case
name = sys.argv[1] if len(sys.argv) > 1 else "unet_default_3c_128"
rec = {}
orig_conv = engine.Tape.conv
def conv_dbg(self, srcs, conv_mod, bn_mod=None, slope=1.0, act=0, out_nchw=False):
    out = orig_conv(self, srcs, conv_mod, bn_mod, slope, act, out_nchw)
    rec.setdefault(MODE, []).append((out.t.clone(), None if out.scale is None else out.scale.clone(), None if out.shift is None else out.shift.clone()))
    return out
engine.Tape.conv = conv_dbg
for MODE in ("fp32", "tf32"):
    ab.set_math(MODE)
    net, sd, cfg, x, y, gold = build_case(name)
    net = net.cuda().train(); x = x.cuda()
    if len(sys.argv) > 2 and sys.argv[2] == "grad":
        out = net(x)
    else:
        with torch.no_grad():
            out = net(x)
    torch.cuda.synchronize()
    rec.setdefault(MODE + "_out", out.detach().clone())
print("final rel", float((rec["fp32_out"] - rec["tf32_out"]).abs().max() / rec["fp32_out"].abs().max()))
for i, (a, b) in enumerate(zip(rec["fp32"], rec["tf32"])):
    ra = float((a[0] - b[0]).abs().max() / a[0].abs().max())
    rs = float((a[1] - b[1]).abs().max() / a[1].abs().max()) if a[1] is not None else -1
    rh = float((a[2] - b[2]).abs().max() / (a[2].abs().max() + 1e-9)) if a[2] is not None else -1
    print(f"layer {i:2d} shape {tuple(a[0].shape)} a_rel {ra:.2e} scale_rel {rs:.2e} shift_rel {rh:.2e}")
