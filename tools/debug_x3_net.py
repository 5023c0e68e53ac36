"""Which part of the tf32x3 backward deviates?  Per-parameter gradient error of the smoke network
vs the fp32 (exact FFMA) mode on the same GPU, with the weight-gradient kernel switched."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
import torch
import atomai_b200 as ab
from atomai_b200.losses_metrics import select_loss
from atomai_b200.nets import Unet

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = Unet(nb_classes=3)
for mod in net.modules():
    if isinstance(mod, torch.nn.BatchNorm2d):
        mod.weight.data.uniform_(0.5, 1.5)
        mod.bias.data.uniform_(-0.2, 0.2)
sd = OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items())
x = torch.rand(2, 1, 64, 64)
y = torch.randint(0, 3, (2, 64, 64))


def grads(math, wgrad_tc=True):
    ab.set_math(math, wgrad_tc=wgrad_tc)
    net.load_state_dict(sd)
    nd = net.to(dev).train()
    nd.zero_grad()
    loss = select_loss("ce", 3)(nd(x.to(dev)), y.to(dev))
    loss.backward()
    g = OrderedDict((k, p.grad.detach().double().cpu()) for k, p in nd.named_parameters())
    net.cpu()
    return g


ref = grads("fp32")
for name, kw in (("x3", dict(math="tf32x3")), ("x3 simt-wgrad", dict(math="tf32x3", wgrad_tc=False)),
                 ("tf32", dict(math="tf32"))):
    g = grads(**kw)
    tot_e = sum(float((g[k] - ref[k]).pow(2).sum()) for k in g)
    tot_r = sum(float(ref[k].pow(2).sum()) for k in g)
    print(f"== {name}: total grel {(tot_e / tot_r) ** 0.5:.2e}")
    for k in g:
        e = float((g[k] - ref[k]).norm() / (ref[k].norm() + 1e-30))
        if e > 1e-4:
            print(f"   {k:32s} {e:.2e}  |ref| {float(ref[k].norm()):.2e}")
