"""tcgen05 issue-rate probe: SM clocks per M128 x N x K8 TF32 MMA vs N, layout, issuing threads."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from atomai_b200 import _C, ops
out = torch.zeros(8, dtype=torch.int64, device="cuda")
iters = 4096
for layout in (0, 1):
    for N in (16, 32, 64, 96, 128):
        for iss in (1, 2):
            out.zero_()
            _C.check(_C.lib().atomai_b200_umma_rate(N, layout, iss, iters, ops.ptr(out), ops.stream_ptr()))
            torch.cuda.synchronize()
            o = out.cpu().tolist()
            print(f"layout {layout} N {N:3d} issuers {iss}: issue {o[0]/iters:6.1f} clk/MMA, retire {max(o[1], o[3])/ (iters*iss):6.1f} clk/MMA (per MMA overall)", flush=True)
