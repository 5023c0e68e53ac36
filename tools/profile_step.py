"""One Segmentor.fit cycle (train step + test forward) at the bench workload, bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off` captures exactly that cycle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import atomai_b200 as ab
from atomai_b200.models import Segmentor
from bench import synth, BATCH, NB_CLASSES

batch = int(sys.argv[1]) if len(sys.argv) > 1 else BATCH
ab.set_math(sys.argv[2] if len(sys.argv) > 2 else "tf32x3")
X, y = synth(2 * batch, 1, 512)
Xt, yt = synth(batch, 2, 512)
m = Segmentor("Unet", nb_classes=NB_CLASSES, seed=1)
m.compile_trainer((X, y, Xt, yt), loss="ce", training_cycles=16, batch_size=batch, full_epoch=False,
                  memory_alloc=64, plot_training_history=False, sync_host=False, filename="/tmp/prof_model")
for e in range(3):
    m.step(e)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m.step(3)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
