"""Run-to-run determinism of Segmentor.fit (fp32 math): device-resident twice, host-resident twice."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np
import golden_utils as gu
import atomai_b200 as ab
from atomai_b200.models import Segmentor
ab.set_math(sys.argv[1] if len(sys.argv) > 1 else "fp32")
X = gu.images(11, 32, 32, 32); y = gu.labels(12, 32, 32, 32, 3)
Xt = gu.images(12, 16, 32, 32); yt = gu.labels(13, 16, 32, 32, 3)
tmp = tempfile.mkdtemp()
res = {}
for tag, alloc in (("dev1", 4), ("dev2", 4), ("host1", 0), ("host2", 0)):
    m = Segmentor("Unet", nb_classes=3, nb_filters=8)
    m.fit(X, y, Xt, yt, training_cycles=5, batch_size=8, memory_alloc=alloc,
          filename=os.path.join(tmp, tag), plot_training_history=False)
    res[tag] = np.array(m.loss_acc["train_loss"], np.float64)
for a, b in (("dev1", "dev2"), ("host1", "host2"), ("dev1", "host1")):
    print(a, b, "max rel diff", float(np.max(np.abs(res[a] - res[b]) / np.abs(res[a]))), res[a][-1], res[b][-1])
