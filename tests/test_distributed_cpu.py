"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: pair sharding and the flat-bucket gradient
all-reduce.  The kernels themselves need no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_pairs_partition():
    from e2e_slam_b200.distributed import shard_pairs
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 4, 8):
            shards = [shard_pairs(n, r, world) for r in range(world)]
            flat = sorted(i for s in shards for i in s)
            assert flat == list(range(n))                                   # disjoint cover
            assert max(map(len, shards)) - min(map(len, shards)) <= 1        # balanced
    with pytest.raises(ValueError):
        shard_pairs(4, 2, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from e2e_slam_b200.distributed import FlatGradBucket, mean_over_ranks, shard_pairs
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    net[1].weight.requires_grad_(False)                     # refinement mode: "bn" tensors frozen
    net[1].bias.requires_grad_(False)
    # the last layer never takes part in the loss -> its grads stay None (like the reference's unused heads)
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(100))
    mine = shard_pairs(8, rank, world)
    loss = net[2](net[1](net[0](x[mine]))).pow(2).mean()
    loss.backward()
    local = {n: (p.grad.clone() if p.grad is not None else None) for n, p in net.named_parameters()}
    bucket = FlatGradBucket(net.parameters(), device="cpu")
    bucket.start().finish()
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    ok = True
    for n, p in net.named_parameters():
        if not p.requires_grad:
            ok &= p.grad is None
            continue
        parts = [g[n] if g[n] is not None else torch.zeros_like(p) for g in gathered]
        ok &= torch.allclose(p.grad, sum(parts) / world, atol=1e-7)
    m = mean_over_ranks(loss)
    losses = [None] * world
    dist.all_gather_object(losses, float(loss))
    ok &= abs(float(m) - sum(losses) / world) < 1e-6
    ok &= bucket.numel == sum(p.numel() for p in net.parameters() if p.requires_grad)
    # second step with the gradients living INSIDE the bucket (adopt_grads): backward accumulates into the views, the
    # all-reduce moves no other bytes, and .grad must still alias the flat buffer afterwards
    bucket.adopt_grads()
    for p_ in net.parameters():
        if p_.grad is not None:
            p_.grad.zero_()
    x2 = torch.randn(8, 6, generator=torch.Generator().manual_seed(200))
    net[2](net[1](net[0](x2[mine]))).pow(2).mean().backward()
    local2 = {n: p_.grad.clone() for n, p_ in net.named_parameters() if p_.requires_grad}
    bucket.start().finish()
    gathered2 = [None] * world
    dist.all_gather_object(gathered2, local2)
    lo, hi = bucket.flat.data_ptr(), bucket.flat.data_ptr() + bucket.flat.numel() * 4
    for n, p_ in net.named_parameters():
        if not p_.requires_grad:
            continue
        ok &= lo <= p_.grad.data_ptr() < hi
        ok &= torch.allclose(p_.grad, sum(g[n] for g in gathered2) / world, atol=1e-7)
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() is True


def test_bucket_without_nvls_support_falls_back():
    """nvls=True on a device without NVLS (here: the CPU) must leave the ordinary bucket in place, not fail."""
    from e2e_slam_b200.distributed import FlatGradBucket
    p = torch.nn.Parameter(torch.zeros(10))
    b = FlatGradBucket([p], device="cpu", nvls=True)
    assert not b.uses_nvls and b.flat.shape == (10,) and b.flat.device.type == "cpu"
