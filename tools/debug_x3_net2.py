"""In-situ check of every conv_tc call of one tf32x3 training step against the exact FFMA kernel
on the same inputs (hooks ops.prep_weights / ops.conv_fwd)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
import copy
import torch
import atomai_b200 as ab
from atomai_b200 import ops, engine
from atomai_b200.losses_metrics import select_loss
from atomai_b200.nets import Unet

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = Unet(nb_classes=3)
for mod in net.modules():
    if isinstance(mod, torch.nn.BatchNorm2d):
        mod.weight.data.uniform_(0.5, 1.5)
        mod.bias.data.uniform_(-0.2, 0.2)
x = torch.rand(2, 1, 64, 64)
y = torch.randint(0, 3, (2, 64, 64))

last = {}
orig_prep, orig_fwd = ops.prep_weights, ops.conv_fwd


def prep(w, mode, math):
    out = orig_prep(w, mode, math)
    last["w"], last["mode"], last["blob"] = w, mode, out
    return out


def fwd(d, wp, bias, out, stats=None):
    orig_fwd(d, wp, bias, out, stats)
    if d.math == ops.MATH_FP32 or wp is not last.get("blob"):
        return
    m0 = d.math
    d.math = ops.MATH_FP32
    d2 = d
    ref = torch.empty_like(out) if out.is_contiguous() else torch.empty(out.shape, device=out.device)
    st2 = torch.zeros_like(stats) if stats is not None else None
    orig_fwd(d2, orig_prep(last["w"], last["mode"], ops.MATH_FP32), bias, ref, st2)
    d.math = m0
    torch.cuda.synchronize()
    e = float((out.double() - ref.double()).norm() / (ref.double().norm() + 1e-30))
    es = float((stats - st2).abs().max() / (st2.abs().max() + 1e-30)) if stats is not None else 0.0
    ctot = sum(d.src[i].C for i in range(d.nsrc))
    print(f"{'dgrad' if last['mode'] == ops.WMODE_DGRAD else 'fwd  '} {d.N}x{d.H}x{d.W} {ctot}->{d.Cout} "
          f"k{d.ks_h} nsrc {d.nsrc} pool {d.src[0].pool} aff {bool(d.src[0].scale)}  rel {e:.1e}  stats {es:.1e}",
          flush=True)


ops.prep_weights, ops.conv_fwd = prep, fwd
ab.set_math("tf32x3")
nd = net.to(dev).train()
nd.zero_grad()
loss = select_loss("ce", 3)(nd(x.to(dev)), y.to(dev))
loss.backward()
