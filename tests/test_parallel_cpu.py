"""world_size-2 gloo tests of the data-parallel host logic (no GPU): the gradient all-reduce gives
every rank the mean gradient, equal to the gradient of the mean loss over the union of shards."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from fsv import parallel
    r, w, _ = parallel.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    if rank == 1:   # ranks start different; broadcast_state must fix that
        for p in net.parameters():
            p.data.add_(1.0)
    parallel.broadcast_state(net)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 6, generator=g)
    lo, hi = parallel.shard_batch(8, rank, world)
    unused = torch.nn.Parameter(torch.zeros(3))          # a parameter without grad on this rank
    # in-place buckets fired from hooks / collect-after-backward / collect with buckets fired from hooks as they complete
    for overlap, stream_fire in ((True, False), (False, False), (False, True)):
        for p in list(net.parameters()) + [unused]:
            p.grad = None
        sync = parallel.GradSync(list(net.parameters()) + [unused], bucket_mb=5e-5, overlap=overlap, stream_fire=stream_fire)      # tiny buckets: several of them
        assert len(sync.buckets) >= 3
        for it in range(2):                               # twice: zero / arm / backward / sync protocol, views stay bound
            sync.zero()
            sync.arm()
            loss = net(x[lo:hi]).square().mean()
            loss.backward()
            assert any(sync.fired) == (overlap or stream_fire)       # buckets were launched from inside backward
            sync()
            assert all(p.grad is sync.views[p] for p in sync.params)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    ref.load_state_dict(net.state_dict())
    ref(x).square().mean().backward()
    err = max(float((a.grad - b.grad).abs().max()) for a, b in zip(net.parameters(), ref.parameters()))
    out[rank] = err
    assert float(unused.grad.abs().max()) == 0.0
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == 2 and all(v < 1e-6 for v in out.values()), dict(out)


def test_shard_batch():
    from fsv import parallel
    assert parallel.shard_batch(16, 3, 8) == (6, 8)
