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

    def __init__(self, group=None, sync_bn: bool = True, p2p: Optional[bool] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.sync_bn = sync_bn
        self.p2p = None               # NVLink peer-memory mailboxes for the SyncBN statistics
        if p2p is None:
            p2p = os.environ.get("ATOMAI_B200_P2P", "1") != "0"
        if p2p and self.world > 1 and sync_bn and torch.cuda.is_available() \
                and dist.get_backend(group) == "nccl":
            try:
                global _MAILBOX          # one set of mapped buffers per process (IPC maps are not re-openable)
                if _MAILBOX is None or _MAILBOX.comm.world != self.world:
                    _MAILBOX = _P2PMailbox(self)
                self.p2p = _MAILBOX
            except Exception as e:  # noqa  (no peer access / IPC: NCCL carries the statistics)
                if self.rank == 0:
                    print(f"atomai_b200: NVLink peer mailboxes unavailable ({type(e).__name__}: {e}); "
                          "SyncBN statistics use NCCL")
                self.p2p = None
            ok = torch.tensor([1 if self.p2p is not None else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.p2p = None

    # --- used by the native tape (engine.Tape.conv / _conv_bwd) for SyncBN
    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1 and self.sync_bn:
            if self.p2p is not None and t.dtype == torch.float64 and t.numel() <= 1024 \
                    and t.is_contiguous():
                self.p2p.allreduce_(t)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def bn_finalize_fused(self, stats, count, gamma, beta, rmean, rvar, momentum, eps, scale, shift,
                          mean, invstd) -> bool:
        """Statistics all-reduce + BatchNorm finalisation as ONE kernel over peer memory; returns
        False when the mailboxes are not available (caller then uses all-reduce + bn_finalize)."""
        if self.p2p is None or self.world == 1 or not self.sync_bn or stats.numel() > 1024:
            return False
        self.p2p.bn_finalize(stats, count, gamma, beta, rmean, rvar, momentum, eps, scale, shift,
                             mean, invstd)
        return True

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


_MAILBOX = None


class _P2PMailbox:
    """Per-rank mailbox + flag buffers mapped into every peer (CUDA IPC through torch's shared
    CUDA storages) for csrc/p2p.cu: small all-reduces as remote stores over NVLink instead of NCCL
    launches.  All ranks must issue the same sequence of calls (they do: same network)."""

    def __init__(self, comm: "Comm"):
        import ctypes as C
        from ._C import check, lib
        self.comm = comm
        w, r = comm.world, comm.rank
        dev = torch.device("cuda", torch.cuda.current_device())
        nd = lib().atomai_b200_p2p_data_bytes(w) // 8
        nf = lib().atomai_b200_p2p_flag_bytes(w) // 8
        # one allocation: [mailbox doubles | flags]
        self.buf = torch.zeros(nd + nf, device=dev, dtype=torch.float64)
        torch.cuda.synchronize()
        handle = (C.c_ubyte * 64)()
        off = C.c_int64(0)
        check(lib().atomai_b200_ipc_export(self.buf.data_ptr(), handle, C.byref(off)))
        mine = (bytes(handle), int(off.value))
        allh = [None] * w
        dist.all_gather_object(allh, mine, group=comm.group)
        dptr, fptr = [], []
        for q in range(w):
            if q == r:
                base = self.buf.data_ptr()
            else:
                hb = (C.c_ubyte * 64).from_buffer_copy(allh[q][0])
                out = C.c_void_p()
                check(lib().atomai_b200_ipc_import(hb, allh[q][1], C.byref(out)))
                base = out.value
            dptr.append(base)
            fptr.append(base + nd * 8)
        arr = C.c_void_p * w
        self.dptr, self.fptr = arr(*dptr), arr(*fptr)
        self.epoch = 0
        torch.cuda.synchronize()
        dist.barrier(group=comm.group)

    def allreduce_(self, t: torch.Tensor) -> None:
        from ._C import check, lib, stream_ptr
        self.epoch += 1
        check(lib().atomai_b200_p2p_allreduce(self.dptr, self.fptr, self.comm.world, self.comm.rank,
                                              self.epoch, t.data_ptr(), t.numel(), stream_ptr()))

    def bn_finalize(self, stats, count, gamma, beta, rmean, rvar, momentum, eps, scale, shift, mean,
                    invstd) -> None:
        from ._C import check, lib, ptr, stream_ptr
        self.epoch += 1
        check(lib().atomai_b200_p2p_bn_finalize(
            self.dptr, self.fptr, self.comm.world, self.comm.rank, self.epoch, ptr(stats),
            scale.numel(), float(count), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar),
            float(momentum), float(eps), ptr(scale), ptr(shift), ptr(mean), ptr(invstd),
            stream_ptr()))


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
