"""N-rank SyncBN data-parallel step == 1-rank step at the global batch (fp32 math).
Launch:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_equivalence.py
Rank 0 prints one JSON line {"loss_rel": ..., "grad_rel": ..., "bn_grad_rel": ..., "world": N}."""
import json
import os
import sys
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_utils as gu  # noqa: E402
import atomai_b200 as ab  # noqa: E402
from atomai_b200.losses_metrics import select_loss  # noqa: E402
from atomai_b200.nets import Unet  # noqa: E402
from atomai_b200.parallel import Comm, init_distributed  # noqa: E402


def main():
    init_distributed()
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    ab.set_math(os.environ.get("DP_MATH", "fp32"))
    B, hw = 4 * world, 64
    x = torch.from_numpy(gu.images(11, B, hw, hw))[:, None].to(dev)
    y = torch.from_numpy(gu.labels(12, B, hw, hw, 3)).to(dev)
    crit = select_loss("ce", 3)

    def make():
        net = Unet(nb_classes=3)
        shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
        net.load_state_dict({k: torch.from_numpy(v) for k, v in gu.fill_state_dict(shapes, 77).items()})
        return net.to(dev).train()

    # data parallel: every rank its slice, SyncBN statistics, gradients summed then / world
    net = make()
    net._comm = Comm(sync_bn=True)
    k = B // world
    loss = crit(net(x[rank * k:(rank + 1) * k]), y[rank * k:(rank + 1) * k])
    loss.backward()
    ld = loss.detach().clone()
    dist.all_reduce(ld)
    ld /= world
    grads = OrderedDict()
    for n_, p in net.named_parameters():
        g = p.grad.detach().clone()
        dist.all_reduce(g)
        grads[n_] = g / world
    rm = net.c1.block[2].running_mean.detach().clone()
    # single process at the global batch (every rank computes it; rank 0 reports)
    ref = make()
    lr = crit(ref(x), y)
    lr.backward()
    te = tr = be = br = 0.0
    worst = []
    for n_, p in ref.named_parameters():
        e = float((grads[n_].double() - p.grad.double()).pow(2).sum())
        r = float(p.grad.double().pow(2).sum())
        worst.append(((e / (r + 1e-300)) ** 0.5, n_))
        te, tr = te + e, tr + r
        if ".2." in n_ or ".5." in n_ or ".8." in n_:       # BatchNorm weight / bias
            be, br = be + e, br + r
    tail = [a for a, b in worst if b.startswith(("c5.", "c6.", "px."))]
    out = {"world": world, "loss_rel": abs(float(ld) - float(lr)) / abs(float(lr)),
           "tail_grad_rel_max": max(tail),
           "grad_rel": (te / tr) ** 0.5, "bn_grad_rel": (be / br) ** 0.5,
           "running_mean_maxdiff": float((rm - ref.c1.block[2].running_mean).abs().max()),
           "worst": [(round(a, 6), b) for a, b in sorted(worst, reverse=True)[:6]],
           "best": [(round(a, 9), b) for a, b in sorted(worst)[:4]],
           "p2p": net._comm.p2p is not None}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
