"""CPU, world_size 2, gloo: the data-parallel contract of the training path.

The biHomE loss is a SUM over the local batch and DDP AVERAGES gradients, so N ranks x B samples must give
(1/N) x the gradient of one process on the N*B samples (BatchNorm-free model: batch statistics are per-rank in the
reference too).  Also checks that rank-seeded pair loaders would draw disjoint streams.  The loss here is the CPU
oracle (the CUDA kernels need a GPU); the all-reduce plumbing is what is under test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class TinyBackbone(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Conv2d(2, 4, 5, stride=4, padding=2), torch.nn.ReLU(),
                                       torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(4, 8))

    def forward(self, p1, p2):
        return 4.0 * torch.tanh(self.net(torch.cat([p1, p2], 1))).reshape(-1, 4, 2)


def oracle_loss(model, p1, p2):
    from oracle import ref_path as R
    from oracle.make_golden import TinyExtractor
    d12, d21 = model(p1, p2), model(p2, p1)
    loss, _ = R.head_double_line(p1, p2, d12, d21, TinyExtractor(), 0.01)
    return loss


def make_batch(n):
    g = torch.Generator().manual_seed(3)
    lo = torch.rand(n, 1, 9, 9, generator=g)
    p1 = torch.nn.functional.interpolate(lo, size=(32, 32), mode='bicubic', align_corners=True)
    return p1, torch.roll(p1, (1, 2), (2, 3))


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    model = TinyBackbone()
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    p1, p2 = make_batch(4 * world)
    sl = slice(4 * rank, 4 * rank + 4)
    loss = oracle_loss(ddp, p1[sl], p2[sl])
    loss.backward()
    total = loss.detach().clone()
    dist.all_reduce(total)
    if rank == 0:
        torch.save({'grads': [p.grad.clone() for p in model.parameters()], 'loss': total}, out)
    dist.destroy_process_group()


def test_two_ranks_match_one_process(tmp_path):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / 'ddp.pt')
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = TinyBackbone()
    p1, p2 = make_batch(8)
    loss = oracle_loss(model, p1, p2)
    loss.backward()
    assert torch.allclose(got['loss'], loss.detach(), rtol=1e-5)
    for g, p in zip(got['grads'], model.parameters()):
        assert torch.allclose(g, p.grad / 2, rtol=1e-4, atol=1e-6)      # DDP mean == (sum over ranks) / world


def test_rank_seeded_loaders_use_disjoint_streams():
    from bihome_b200.data.gpu_pairs import rank_seed
    assert len({rank_seed(42, r) for r in range(8)}) == 8
    assert rank_seed(42, 0) != rank_seed(43, 0)


class BnNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(1, 3, 3, padding=1)
        self.bn = torch.nn.BatchNorm2d(3)

    def forward(self, x):
        return self.bn(self.conv(x)).square().mean()


def bn_worker(rank, world, port, out):
    from bihome_b200 import engine
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    model = BnNet()
    ddp = engine.data_parallel(model)              # the wrapper train.py and bench.py use
    x = torch.randn(6, 1, 8, 8, generator=torch.Generator().manual_seed(10 + rank)) + 3.0 * rank
    for _ in range(2):                             # the second forward is where broadcast_buffers=True would overwrite
        ddp(x).backward()
    torch.save({'mean': model.bn.running_mean.clone(), 'var': model.bn.running_var.clone(),
                'grad': model.conv.weight.grad.clone()}, out + str(rank))
    dist.destroy_process_group()


def test_batchnorm_statistics_stay_rank_local(tmp_path):
    """no SyncBN and no buffer broadcast -- every rank keeps the running statistics of ITS batches
    (the reference's per-GPU BatchNorm), while gradients are still averaged."""
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / 'bn.pt')
    mp.spawn(bn_worker, args=(2, port, out), nprocs=2, join=True)
    got = [torch.load(out + str(r)) for r in range(2)]
    assert not torch.allclose(got[0]['mean'], got[1]['mean'], atol=1e-3)
    assert torch.allclose(got[0]['grad'], got[1]['grad'])
    for rank in range(2):
        torch.manual_seed(0)
        model = BnNet()
        x = torch.randn(6, 1, 8, 8, generator=torch.Generator().manual_seed(10 + rank)) + 3.0 * rank
        for _ in range(2):
            model(x)
        assert torch.allclose(got[rank]['mean'], model.bn.running_mean, atol=1e-6)
        assert torch.allclose(got[rank]['var'], model.bn.running_var, atol=1e-6)


def flat_worker(rank, world, port, out):
    """the data-parallel path of engine.GraphedStep without a DDP wrapper: parameters broadcast once, gradients accumulated
    into views of one flat buffer, ONE all-reduce (mean) of that buffer -- must equal what DDP computes"""
    from bihome_b200 import engine
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(rank)                       # different initial weights per rank: sync_module_state must fix that
    model = TinyBackbone().to(memory_format=torch.channels_last)
    engine.sync_module_state(model)
    params = [p for p in model.parameters() if p.requires_grad]
    flat = engine.flatten_gradients(params)
    assert all(p.grad.stride() == p.stride() and p.grad.shape == p.shape for p in params)
    p1, p2 = make_batch(4 * world)
    sl = slice(4 * rank, 4 * rank + 4)
    results = []
    for rep in range(2):                          # the second pass checks zero-and-accumulate on the same buffer
        flat.zero_()
        oracle_loss(model, p1[sl], p2[sl]).backward()
        engine.average_gradients(params, flat)
        results.append([p.grad.clone() for p in params])
    # coalesced path (gradients that are not views of one buffer)
    for p in params:
        p.grad = None
    oracle_loss(model, p1[sl], p2[sl]).backward()
    engine.average_gradients(params)
    if rank == 0:
        torch.save({'flat': results, 'coalesced': [p.grad.clone() for p in params],
                    'weights': [p.detach().clone() for p in params]}, out)
    dist.destroy_process_group()


def test_flat_gradient_all_reduce_matches_ddp(tmp_path):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / 'flat.pt')
    mp.spawn(flat_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)                          # rank 0's weights are the ones every rank ends up with
    ref = TinyBackbone()
    for w, p in zip(got['weights'], ref.parameters()):
        assert torch.equal(w, p.detach())
    p1, p2 = make_batch(8)
    oracle_loss(ref, p1, p2).backward()
    for k, p in enumerate(ref.parameters()):
        want = p.grad / 2
        for rep in got['flat']:
            assert torch.allclose(rep[k], want, rtol=1e-4, atol=1e-6), k
        assert torch.allclose(got['coalesced'][k], want, rtol=1e-4, atol=1e-6), k
