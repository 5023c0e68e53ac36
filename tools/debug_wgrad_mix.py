"""Gradient error of the mixed mode (tf32x3 forward/dgrad + tf32 wgrad) vs the exact-fp32 mode on
the same GPU, per parameter, at two problem sizes (the wgrad reduction length grows with N*H*W)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
import torch
import atomai_b200 as ab
from atomai_b200.losses_metrics import select_loss
from atomai_b200.nets import Unet

dev = torch.device("cuda:0")
for (n, hw) in ((2, 64), (4, 256)):
    torch.manual_seed(0)
    net = Unet(nb_classes=3)
    for mod in net.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.uniform_(-0.2, 0.2)
    sd = OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items())
    x = torch.rand(n, 1, hw, hw)
    y = torch.randint(0, 3, (n, hw, hw))

    def grads(**kw):
        ab.set_math(**kw)
        net.load_state_dict(sd)
        nd = net.to(dev).train()
        nd.zero_grad()
        loss = select_loss("ce", 3)(nd(x.to(dev)), y.to(dev))
        loss.backward()
        g = OrderedDict((k, p.grad.detach().double().cpu()) for k, p in nd.named_parameters())
        net.cpu()
        return g

    ref = grads(mode="fp32")
    for name, kw in (("x3", dict(mode="tf32x3")), ("x3+wgrad-tf32", dict(mode="tf32x3", wgrad_math="tf32")),
                     ("tf32", dict(mode="tf32"))):
        g = grads(**kw)
        tot_e = sum(float((g[k] - ref[k]).pow(2).sum()) for k in g)
        tot_r = sum(float(ref[k].pow(2).sum()) for k in g)
        worst = max(((float((g[k] - ref[k]).norm() / (ref[k].norm() + 1e-30)), k) for k in g))
        print(f"{n}x{hw}^2 {name:14s}: total grel {(tot_e / tot_r) ** 0.5:.2e}   worst param {worst[0]:.2e} ({worst[1]})")
ab.set_math("tf32x3")
