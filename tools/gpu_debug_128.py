import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import atomai_b200 as ab
from test_unet_gpu import build_case, rel, logits_view, BIG, grad_errors
from atomai_b200.losses_metrics import select_loss
cuda = torch.device("cuda")
for math in ("fp32", "tf32"):
    for order in ("train_only", "eval_then_train", "train_twice"):
        ab.set_math(math)
        net, sd, cfg, x, y, gold = build_case(BIG)
        net = net.to(cuda); x = x.to(cuda); y = y.to(cuda)
        scale = float(gold["logits_absmax"]) if "logits_absmax" in gold.files else None
        if order == "eval_then_train":
            net.eval()
            with torch.no_grad():
                le = net(x)
            print(math, order, "eval rel", rel(logits_view(le.cpu().numpy(), gold), gold["logits_eval"], scale))
        net.train(); net.zero_grad()
        lt = net(x)
        print(math, order, "train rel", rel(logits_view(lt.detach().cpu().numpy(), gold), gold["logits_train"], scale), flush=True)
        if order == "train_twice":
            lt = net(x)
            print(math, order, "train rel 2", rel(logits_view(lt.detach().cpu().numpy(), gold), gold["logits_train"], scale), flush=True)
        loss = select_loss("ce", cfg["nb_classes"])(lt, y)
        loss.backward()
        grel, per, gnorm = grad_errors(net, gold)
        print(math, order, "loss", loss.item(), float(gold["loss_train"]), "grel", grel, flush=True)
