"""
Data parallelism for the trainers: one process per GPU (torchrun), torch.distributed/NCCL over
NVLink 5 / NVSwitch as the plumbing.  The reference has no parallelism at all;
this adds what §8e specifies:

  * every mini-batch is split along N (rank r keeps rows [r*B/G, (r+1)*B/G));
  * ONE all-reduce per step of a flat, pre-allocated gradient bucket (2.4 MB for the default
    Unet: latency- not bandwidth-bound, so a single bucket beats per-tensor reduction);
  * optional synchronised BatchNorm statistics (2*C doubles per BN layer) so that the result
    equals the single-process reference at the global batch size;
  * the 1/world averaging is folded into the fused Adam kernel (grad_scale).
"""
import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class Comm:
    """Thin wrapper over a process group; also usable with gloo on CPU for the host-logic tests."""

    def __init__(self, group=None, sync_bn: bool = True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.sync_bn = sync_bn

    # --- used by the native tape (engine.Tape.conv / _conv_bwd) for SyncBN
    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1 and self.sync_bn:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_count(self, count: int) -> int:
        return count * self.world if (self.world > 1 and self.sync_bn) else count

    # --- used by the trainers
    def allreduce_any_(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def broadcast_(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.world > 1:
            dist.broadcast(t, src=src, group=self.group)
        return t


class GradBucket:
    """Flat fp32 gradient buffer: param.grad tensors are views into it, so backward writes land in
    the bucket and one all-reduce covers the whole model."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        self.views = []
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(v)
            off += p.numel()
        self.post_scale = 1.0     # 1/world for optimizers that do not fold the averaging in

    def attach(self) -> None:
        """(Re)point every param.grad at its slice of the bucket (call after zero_grad())."""
        for p, v in zip(self.params, self.views):
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v

    def zero_(self) -> None:
        self.flat.zero_()

    def allreduce(self, comm: Comm) -> None:
        comm.allreduce_any_(self.flat)
        if self.post_scale != 1.0:
            self.flat.mul_(self.post_scale)


def init_distributed(backend: Optional[str] = None) -> Comm:
    """Initialise torch.distributed from the torchrun environment (no-op when WORLD_SIZE <= 1)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend)
    return Comm()


def broadcast_model(model: torch.nn.Module, comm: Comm) -> None:
    """Make every rank start from rank 0's parameters and buffers."""
    if comm.world == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        comm.broadcast_(t.data)
