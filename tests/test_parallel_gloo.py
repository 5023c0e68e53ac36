"""Data-parallel host logic on CPU with gloo, world_size 2: sharding, flat gradient bucket
all-reduce, SyncBN statistic reduction plumbing (atomai_b200/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from atomai_b200.parallel import Comm, GradBucket, broadcast_model, init_distributed
    from atomai_b200.utils.preproc import shard_batches
    comm = init_distributed("gloo")
    assert comm.world == world and comm.rank == rank
    torch.manual_seed(100 + rank)
    lin = torch.nn.Linear(4, 3)
    broadcast_model(lin, comm)
    w0 = lin.weight.detach().clone()
    gathered = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(gathered, w0)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    # gradient bucket: grads become views of one flat buffer; one all-reduce sums them
    bucket = GradBucket(lin.parameters())
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    bucket.attach()
    assert lin.weight.grad.data_ptr() == bucket.flat.data_ptr()
    bucket.allreduce(comm)
    assert torch.all(bucket.flat == sum(range(1, world + 1)))
    assert torch.all(lin.bias.grad == sum(range(1, world + 1)))
    # SyncBN statistics (sum, sumsq, count)
    st = torch.tensor([1.0 + rank, 2.0], dtype=torch.float64)
    comm.allreduce_sum_(st)
    assert st.tolist() == [3.0, 4.0] and comm.allreduce_count(10) == 20
    nosync = Comm(sync_bn=False)
    st2 = torch.tensor([1.0], dtype=torch.float64)
    nosync.allreduce_sum_(st2)
    assert st2.item() == 1.0 and nosync.allreduce_count(10) == 10
    # sharding: union over ranks is the global batch, in order
    batches = [torch.arange(8).reshape(8, 1), torch.arange(8, 16).reshape(8, 1)]
    mine = shard_batches(batches, rank, world)
    allb = [torch.zeros(4, 1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allb, mine[1])
    assert torch.cat(allb).flatten().tolist() == list(range(8, 16))
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = True


def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert len(ret) == world
