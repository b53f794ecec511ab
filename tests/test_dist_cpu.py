"""World-size-2 gloo test (CPU) of the data-parallel host logic: frame sharding and the flat-bucket gradient
all-reduce that replaces DistributedDataParallel on the hot path (tools/stage1_cutmix_train.py L142)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from toda_b200.dist import FlatGradBucket, shard_frames


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    bucket = FlatGradBucket(net.parameters())
    for step in range(2):
        bucket.zero()
        assert all(p.grad is None for p in net.parameters())
        x = torch.full((4, 5), float(rank + 1 + step))
        net(x).sum().backward()
        local = torch.cat([p.grad.flatten() for p in net.parameters()]).clone()
        bucket.all_reduce_mean()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(bucket.flat, sum(gathered) / world, rtol=1e-6, atol=1e-7)
        # the averaged gradients are back in every parameter's .grad
        assert torch.equal(torch.cat([p.grad.flatten() for p in net.parameters()]), bucket.flat)
    first = shard_frames(100, 4, rank)
    out[rank] = (first, float(bucket.flat.sum()))
    dist.barrier()
    dist.destroy_process_group()


def _worker_unused(rank, world, port, out):
    """One parameter gets no gradient on rank 1 only (a frozen branch / an empty shard): every replica must still end
    up with the same averaged gradient for it, and the collectives must be issued in the same order on both ranks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    a, b, c = torch.nn.Linear(5, 7), torch.nn.Linear(7, 3), torch.nn.Linear(5, 3)
    params = list(a.parameters()) + list(b.parameters()) + list(c.parameters())
    bucket = FlatGradBucket(params, bucket_bytes=64)           # several small buckets
    assert len(bucket.buckets) >= 3
    bucket.zero()
    x = torch.full((4, 5), float(rank + 1))
    y = b(a(x)).sum()
    if rank == 0:
        y = y + c(x).sum()                                      # `c` is used on rank 0 only
    y.backward()
    had = [p.grad is not None for p in params]
    local = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten() for p in params]).clone()
    bucket.all_reduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(bucket.flat, sum(gathered) / world, rtol=1e-6, atol=1e-7)
    assert all(p.grad is not None for p in params)              # also where this rank had produced none
    assert torch.equal(torch.cat([p.grad.flatten() for p in params]), bucket.flat)
    out[rank] = (had, float(bucket.flat.abs().sum()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_parameter_unused_on_one_rank():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_unused, args=(world, port, out), nprocs=world, join=True)
    assert all(out[0][0]) and not all(out[1][0])               # rank 1 really lacked gradients for `c`
    assert abs(out[0][1] - out[1][1]) < 1e-6 and out[0][1] > 0


def test_flat_bucket_allreduce_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] == 100 and out[1][0] == 104            # disjoint frame ranges per rank
    assert abs(out[0][1] - out[1][1]) < 1e-6                # identical averaged gradients on both ranks


def _worker_no_sync(rank, world, port, out):
    """A backward pass that only rank 0 runs (bench.py's per-kernel roofline pass) must not launch a collective: inside
    no_sync() the gradient hooks stay quiet, and the next synchronised step still pairs up on both ranks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    bucket = FlatGradBucket(net.parameters(), bucket_bytes=64)
    if rank == 0:
        with bucket.no_sync():
            bucket.zero()
            net(torch.ones(4, 5)).sum().backward()
            bucket.all_reduce_mean()             # a no-op inside no_sync
            assert not bucket._work
    bucket.zero()
    x = torch.full((4, 5), float(rank + 1))
    net(x).sum().backward()
    local = torch.cat([p.grad.flatten() for p in net.parameters()]).clone()
    bucket.all_reduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(bucket.flat, sum(gathered) / world, rtol=1e-6, atol=1e-7)
    out[rank] = float(bucket.flat.sum())
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_no_sync_pass_on_one_rank():
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker_no_sync, args=(world, port, out), nprocs=world, join=True)
    assert abs(out[0] - out[1]) < 1e-6
