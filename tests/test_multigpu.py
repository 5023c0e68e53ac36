"""Hardware data-parallel equivalence (needs >= 2 GPUs; the 1-GPU round-end run skips it, the
recorded 2- and 8-GPU runs are in profiles/r02_scaling_and_dp_equivalence.md): an N-rank SyncBN step equals the 1-rank step
at the global batch — loss, every gradient (incl. BatchNorm gamma/beta) and the running statistics."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_syncbn_equals_single_rank(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        "29731", os.path.join(ROOT, "tools", "dp_equivalence.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["world"] == 2
    # Forward quantities and the gradients of the last blocks agree to fp32 rounding.  The
    # gradients of the layers further upstream are an ill-conditioned function of the forward
    # (BatchNorm backward over few samples: ANY 1e-7 perturbation — another summation order, the
    # tf32x3 mode, a second rank — moves them by 3-5e-3, see tools/debug_x3_net.py and
    # profiles/r02_scaling_and_dp_equivalence.md), so the whole-model bound is 1e-2.
    assert out["loss_rel"] <= 1e-6, out
    assert out["running_mean_maxdiff"] <= 1e-6, out
    print(json.dumps(out))
    assert out["tail_grad_rel_max"] <= 2e-2, out      # c5 / c6 / px parameters (8 ranks: 5.6e-4)
    assert out["grad_rel"] <= 2e-2 and out["bn_grad_rel"] <= 2e-2, out
