"""CPU (gloo, world_size 2): the data-parallel gradient averaging used by the N>1 training path (sat_b200/dist.py)
gives every rank the mean of the per-rank gradients, for parameters re-homed into flat buckets."""
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


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import sat_b200  # noqa: F401
    from sat_b200.dist import FlatGradBuckets
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s)) for s in [(7, 5), (11,), (3, 4, 2), (1,)]]
    buckets = FlatGradBuckets(params, bucket_bytes=200)          # forces several buckets
    assert len(buckets.buckets) >= 2
    buckets.zero()
    # a "backward": autograd accumulates into the bucket views in place
    loss = sum(((rank + 1) * (i + 1)) * p.sum() for i, p in enumerate(params))
    loss.backward()
    params[1].grad = params[1].grad.clone()                       # simulate a re-assigned .grad (rebind must catch it)
    buckets.allreduce_mean(world)
    ok = True
    for i, p in enumerate(params):
        expect = sum((r + 1) * (i + 1) for r in range(world)) / world
        ok = ok and torch.allclose(p.grad, torch.full_like(p, expect))
        ok = ok and any(f.data_ptr() <= p.grad.data_ptr() < f.data_ptr() + f.numel() * 4 for f in buckets.flat)
    out[rank] = ok
    dist.destroy_process_group()


def test_flat_bucket_allreduce_mean_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _worker_overlap(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import sat_b200  # noqa: F401
    from sat_b200.dist import OverlappedGradReducer
    torch.manual_seed(0)
    dec = [torch.nn.Parameter(torch.randn(s)) for s in [(7, 5), (11,)]]
    enc = [torch.nn.Parameter(torch.randn(s)) for s in [(3, 4, 2), (1,), (6, 6), (9,)]]
    unused = torch.nn.Parameter(torch.randn(4))                  # never receives a gradient: its bucket is left to finish()
    red = OverlappedGradReducer(dec, enc + [unused], bucket_bytes=100)
    assert len(red.buckets) >= 3 and red.buckets[0] == dec
    launched = []
    orig = red._launch
    red._launch = lambda bi: (launched.append(bi), orig(bi))[1]
    ok = True
    for step in range(2):                                         # two steps: buckets are re-zeroed, hooks re-armed
        red.prepare()
        params = dec + enc
        loss = sum(((rank + 1) * (i + 1) * (step + 1)) * p.sum() for i, p in enumerate(params))
        loss.backward()
        n_in_backward = len(launched)
        red.finish()
        for i, p in enumerate(params):
            expect = sum((r + 1) * (i + 1) * (step + 1) for r in range(world)) / world
            ok = ok and torch.allclose(p.grad, torch.full_like(p, expect))
        ok = ok and unused.grad is None                             # no staging buffers: a parameter without a gradient stays without
        ok = ok and n_in_backward >= len(red.buckets) - 1          # every bucket but the unused one started inside backward()
        launched.clear()
    red.close()
    out[rank] = ok
    dist.destroy_process_group()


def test_overlapped_grad_reducer_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_overlap, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
