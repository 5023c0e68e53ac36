"""Determinism probe: same-seed Segmentor fits, losses per step, tf32 vs fp32 math."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import golden_utils as gu
import atomai_b200 as ab
from atomai_b200.models import Segmentor
X = gu.images(1, 24, 64, 64); y = gu.labels(2, 24, 64, 64, 3)
Xt = gu.images(5, 8, 64, 64); yt = gu.labels(6, 8, 64, 64, 3)
for math in ("fp32", "tf32"):
    for wg in (True, False):
        ab.set_math(math, wg) if math == "tf32" else ab.set_math(math)
        for rep in range(3):
            m = Segmentor("Unet", nb_classes=3)
            m.fit(X, y, Xt, yt, training_cycles=3, batch_size=8, filename="/tmp/det", plot_training_history=False, print_loss=100)
            print(math, "wgrad_tc" if wg else "wgrad_simt", rep, ["%.6f" % v for v in m.loss_acc["train_loss"]], ["%.6f" % v for v in m.loss_acc["test_loss"]], flush=True)
        if math == "fp32": break
