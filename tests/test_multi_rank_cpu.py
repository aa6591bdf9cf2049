"""CPU, world_size 2, gloo: the data-parallel host logic (flat gradient all-reduce, rank seeding, the bench
`max over ranks` reduction).  The kernels themselves have no collective."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, 'style-big-gan_b200'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sgb200.training import FlatGradAllReduce
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5, dtype=torch.float32) * (rank + 1)
    # params[2] has no grad on any rank (e.g. frozen layer): must be skipped consistently
    n = FlatGradAllReduce(params)()
    ok = n == 17
    ok &= torch.allclose(params[0].grad, torch.full((3, 4), 1.5))
    ok &= torch.allclose(params[1].grad, torch.arange(5, dtype=torch.float32) * 1.5)
    ok &= params[2].grad is None
    # FlatGrads: the gradients gathered into one flat buffer and averaged in place with one all-reduce (the in-graph path of
    # sgb200.training.Trainer); a parameter without a gradient keeps grad None and its slot stays zero
    from sgb200.training import FlatGrads
    ps = [torch.nn.Parameter(torch.ones(3, 4)), torch.nn.Parameter(torch.ones(5)), torch.nn.Parameter(torch.ones(2, 2))]
    fg = FlatGrads(ps)
    fg.flat.fill_(7.0)                                               # stale values from an earlier phase
    ((ps[0] * (rank + 1)).sum() + (ps[1] * ps[1] * (rank + 2)).sum()).backward()
    fg.gather()
    ok &= ps[0].grad.data_ptr() == fg.views[0].data_ptr()           # .grad now aliases the flat buffer
    ok &= fg.all_reduce_mean() == 21
    ok &= torch.allclose(ps[0].grad, torch.full((3, 4), 1.5)) and torch.allclose(ps[1].grad, torch.full((5,), 5.0))
    ok &= ps[2].grad is None and float(fg.flat[17:].abs().sum()) == 0.0
    # bench.py's timing reduction: max over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok &= float(t) == 11.0
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_flat_grad_allreduce_two_ranks():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_single_rank_is_a_noop():
    sys.path.insert(0, os.path.join(ROOT, 'style-big-gan_b200'))
    from sgb200.training import FlatGradAllReduce
    p = torch.nn.Parameter(torch.zeros(2))
    p.grad = torch.ones(2)
    assert FlatGradAllReduce([p])() == 0
    assert torch.equal(p.grad, torch.ones(2))
