"""N>1 host logic on CPU: two `gloo` processes run the view sharding and the flat-gradient all-reduce that
`Distiller.photometric_substep / fusion_substep` use on the GPU box with NCCL (sparsefusion_b200/distillation.py).

What is covered here without a GPU: (1) every rank draws the same per-step permutation and the ranks pick distinct
target views (the reference's single process picks entry 1, distillation.py:264); (2) FlatAdam turns the parameters
and their .grad into views of one flat buffer, so autograd accumulates straight into the buffer that is all-reduced;
(3) after `sync_grads` every rank holds bit-identical summed gradients and the returned grad_scale makes the update
the mean over ranks.  The fused Adam kernel itself is CUDA-only (tests/test_distillation_gpu.py).
"""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sparsefusion_b200.distillation import FlatAdam, shard_target_view


class _TinyField(torch.nn.Module):
    """two parameter groups like NeRFNetwork.get_params (network_grid.py:223-234): encoder lr*10, MLPs lr"""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(7)
        self.embeddings = torch.nn.Parameter(torch.randn(37, 2, generator=g))
        self.w0 = torch.nn.Parameter(torch.randn(8, 4, generator=g))
        self.w1 = torch.nn.Parameter(torch.randn(4, 8, generator=g))

    def get_params(self, lr):
        return [{'params': (p for p in [self.embeddings]), 'lr': lr * 10},
                {'params': (p for p in [self.w0, self.w1]), 'lr': lr}]

    def forward(self, idx, x):
        return (torch.tanh(self.embeddings[idx] @ self.w1[:2, :] @ self.w0) * x).sum()


def _worker(rank, world, store_path, out_dir):
    dist.init_process_group('gloo', init_method=f'file://{store_path}', rank=rank, world_size=world)
    try:
        net = _TinyField()
        opt = FlatAdam(net, lr=5e-4)
        assert opt.flat.numel() == 37 * 2 + 32 + 32 and len(opt.groups) == 2
        assert opt.groups[0] == (0, 74, 5e-3) and opt.groups[1][:2] == (74, 138)
        for p in net.parameters():                                   # parameters and grads are views into the flat buffers
            assert p.data.untyped_storage().data_ptr() == opt.flat.untyped_storage().data_ptr()
            assert p.grad.untyped_storage().data_ptr() == opt.grad.untyped_storage().data_ptr()

        gen = torch.Generator().manual_seed(0)                      # same seed on every rank, as in Distiller.__init__
        views, local, summed = [], [], []
        for step in range(3):
            perm = torch.randperm(5, generator=gen)
            vi = shard_target_view(perm, rank)
            views.append(vi)
            opt.zero_grad()
            loss = net(torch.tensor([vi, vi + 3]), torch.full((4,), 1.0 + vi))
            loss.backward()                                          # accumulates in place into opt.grad
            local.append(opt.grad.clone())
            scale = opt.sync_grads(world, None)
            assert scale == 1.0 / world
            summed.append(opt.grad.clone())
        torch.save(dict(views=views, local=local, summed=summed, perm_last=perm), os.path.join(out_dir, f'r{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_view_shard_and_grad_allreduce():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        store = os.path.join(d, 'store')
        mp.spawn(_worker, args=(world, store, d), nprocs=world, join=True)
        r = [torch.load(os.path.join(d, f'r{k}.pt')) for k in range(world)]
    assert torch.equal(r[0]['perm_last'], r[1]['perm_last'])                     # identical permutation stream
    for step in range(3):
        assert r[0]['views'][step] != r[1]['views'][step]                          # distinct target views per step
        assert not torch.equal(r[0]['local'][step], r[1]['local'][step])
        want = r[0]['local'][step] + r[1]['local'][step]
        assert torch.equal(r[0]['summed'][step], r[1]['summed'][step])             # bit-identical on every rank
        assert torch.allclose(r[0]['summed'][step], want, rtol=0, atol=0)


def test_single_rank_is_the_reference_choice():
    perm = torch.tensor([4, 2, 0, 1, 3])
    assert shard_target_view(perm, 0) == 2                                        # entry 1 (distillation.py:264)
    assert [shard_target_view(perm, k) for k in range(5)] == [2, 0, 1, 3, 4]      # wraps
    net = _TinyField()
    opt = FlatAdam(net)
    opt.grad.fill_(3.0)
    assert opt.sync_grads(1, None) == 1.0 and float(opt.grad[0]) == 3.0            # no collective when world_size == 1
    with pytest.raises(RuntimeError):                                               # the Adam kernel is CUDA-only: loud on CPU
        opt.step()


def _mb_worker(rank, world, store_path, out_dir):
    """host logic of the view-batched step on R gloo ranks: same permutation stream, disjoint shards covering the step's V views, per-view seeds that do
    not depend on the rank, photometric term owned by exactly one rank, and ONE all-reduce of the flat buffer per step"""
    from sparsefusion_b200.distillation import shard_views, step_views, view_seed
    dist.init_process_group('gloo', init_method=f'file://{store_path}', rank=rank, world_size=world)
    try:
        net = _TinyField()
        opt = FlatAdam(net, lr=5e-4)
        gen = torch.Generator().manual_seed(0)
        rec = []
        n_coll = 0
        for itr in range(1001, 1004):
            idx = int(torch.randperm(2, generator=gen)[0])
            perm = torch.randperm(7, generator=gen)
            torch.rand(1, generator=gen)
            views = step_views(perm, 4)
            mine = shard_views(views, rank, world)
            opt.zero_grad()
            if itr % world == rank:                                   # the photometric term: one owner per iteration
                net(torch.tensor([idx, idx + 1]), torch.full((4,), 0.5)).backward()
            for v in mine:                                            # mean over the step's V views
                (net(torch.tensor([v, v + 3]), torch.full((4,), 1.0 + v)) / 4).backward()
            local = opt.grad.clone()
            opt.sync_grads(world, None)
            n_coll += 1
            rec.append(dict(views=views, mine=mine, owner=itr % world == rank, local=local, summed=opt.grad.clone(),
                            seeds=[view_seed(11, itr, v, s) for v in views for s in (0, 1, 2)]))
        torch.save(dict(rec=rec, n_coll=n_coll), os.path.join(out_dir, f'mb{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_view_batched_step_host_logic_two_ranks():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_mb_worker, args=(world, os.path.join(d, 'store'), d), nprocs=world, join=True)
        r = [torch.load(os.path.join(d, f'mb{k}.pt')) for k in range(world)]
    assert r[0]['n_coll'] == r[1]['n_coll'] == 3                                   # one collective per step
    # the single-rank result of the same three steps
    net = _TinyField()
    opt = FlatAdam(net, lr=5e-4)
    for step in range(3):
        a, b = r[0]['rec'][step], r[1]['rec'][step]
        assert a['views'] == b['views'] and len(set(a['views'])) == 4             # same permutation stream on every rank
        assert sorted(a['mine'] + b['mine']) == sorted(a['views']) and not set(a['mine']) & set(b['mine'])
        assert a['seeds'] == b['seeds'] and len(set(a['seeds'])) == len(a['seeds'])   # per-(iteration, view, stream) seeds, independent of the rank
        assert a['owner'] != b['owner']                                             # exactly one rank computes the photometric term
        assert torch.equal(a['summed'], b['summed'])
        assert torch.allclose(a['summed'], a['local'] + b['local'], rtol=0, atol=0)
    from sparsefusion_b200.distillation import KeyedNoise
    n1 = KeyedNoise([5, 9], 'cpu', chunk=4)
    n2 = KeyedNoise([9], 'cpu', chunk=4)
    like2, like1 = torch.empty(2, 3, 2, 2), torch.empty(1, 3, 2, 2)
    for _ in range(6):                                                              # crosses a chunk refill
        assert torch.equal(n1(like2)[1], n2(like1)[0])                              # a view's noise does not depend on the batch it shares
