"""CPU, world_size 2 over gloo: host-side logic of the view-parallel SDS step (pose sharding, shared / per-rank seeds,
flat-bucket gradient all-reduce == single-process accumulation)."""
import importlib
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _fake_grads(pose, n_enc=4096):
    g = torch.Generator().manual_seed(1234 + pose)
    enc = torch.randn(n_enc, generator=g)
    mlp = [torch.randn(s, generator=g) for s in ((64, 32), (64,), (64, 64), (64,), (4, 64), (4,))]
    return enc, mlp


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    par = importlib.import_module("make-it-3d_b200.parallel")
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, l, w = par.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    enc_p = torch.nn.Parameter(torch.zeros(4096))
    mlp_p = [torch.nn.Parameter(torch.zeros(s)) for s in ((64, 32), (64,), (64, 64), (64,), (4, 64), (4,))]
    reducer = par.GradientAllReduce(enc_p, mlp_p, op="sum")
    poses = []
    for step in range(3):
        pose = par.pose_index(step, rank, world)
        poses.append(pose)
        enc_g, mlp_g = _fake_grads(pose)
        enc_p.grad = enc_g.clone()
        for p, g in zip(mlp_p, mlp_g):
            p.grad = g.clone()
        if step == 1:
            mlp_p[3].grad = None                      # a parameter that received no gradient this step
        reducer()
        want_enc = sum(_fake_grads(par.pose_index(step, rr, world))[0] for rr in range(world))
        assert torch.allclose(enc_p.grad, want_enc, atol=1e-6)
        for i, p in enumerate(mlp_p):
            want = sum((_fake_grads(par.pose_index(step, rr, world))[1][i] if not (step == 1 and i == 3) else torch.zeros_like(p))
                       for rr in range(world))
            assert torch.allclose(p.grad, want, atol=1e-6), (step, i)
    assert par.shared_seed(7, 5) == par.shared_seed(7, 5) and par.rank_seed(7, 0) != par.rank_seed(7, 1)
    assert par.max_over_ranks(float(rank + 1), "cpu") == float(world)
    out[rank] = poses
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_sequential_accumulation():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    # union of the per-rank pose sequences == the single-process sequence 0..5
    assert sorted(out[0] + out[1]) == list(range(6)) and out[0] == [0, 2, 4] and out[1] == [1, 3, 5]


def _img(view, HW, C=5):
    """synthetic per-view 'image' both the producer ranks and the owner can compute: value = f(view, pixel, channel)"""
    p = torch.arange(HW, dtype=torch.float32)[:, None]
    c = torch.arange(C, dtype=torch.float32)[None, :]
    return 1000.0 * view + p + 0.01 * c


def _ray_parallel_worker(rank, world, port):
    sys.path.insert(0, ROOT)
    par = importlib.import_module("make-it-3d_b200.parallel")
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    par.init_distributed(backend="gloo")
    rp = par.RayParallel()
    HW, C = 64, 5
    rpv = HW // world
    # forward: this rank rendered pixel i*world + rank of every view v (batch row v*rpv + i)
    mine = torch.arange(rpv) * world + rank
    packed = torch.cat([_img(v, HW, C)[mine] for v in range(world)])
    full = rp.fragments_to_owner(packed, world)
    assert torch.equal(full, _img(rank, HW, C)), "owner must receive its complete row-major image"
    # backward: the owner's gradient image goes back to the ranks that rendered each pixel, in their batch-row order
    gfull = -_img(rank, HW, C)
    frag = rp.owner_to_fragments(gfull, world)
    want = torch.cat([-_img(v, HW, C)[mine] for v in range(world)])
    assert torch.equal(frag, want)
    # round trip is the identity
    assert torch.equal(rp.owner_to_fragments(rp.fragments_to_owner(packed, world), world), packed)
    # per-view sample counts: out[r, v] = rank r's count of view v
    my_counts = torch.tensor([100 * rank + v for v in range(world)], dtype=torch.int32)
    allc = rp.gather_counts(my_counts, torch.empty(world, world, dtype=torch.int32))
    assert allc.tolist() == [[100 * r + v for v in range(world)] for r in range(world)]
    # regulariser sums: every rank holds its share of every view's mean
    part = torch.tensor([[1.0 + rank, 10.0 + rank], [2.0 * (rank + 1), 20.0]])[:, :world]
    tot = rp.reduce_losses(part)
    assert torch.allclose(tot, sum(torch.tensor([[1.0 + r, 10.0 + r], [2.0 * (r + 1), 20.0]])[:, :world] for r in range(world)))
    assert part[0, 0] == 1.0 + rank                      # out of place
    gl = rp.gather_loss_grads(torch.tensor([0.11 * (rank + 1), 1.0]))
    assert torch.allclose(gl, torch.tensor([[0.11 * (r + 1), 1.0] for r in range(world)]))
    dist.barrier()
    dist.destroy_process_group()


def test_ray_parallel_exchange_over_gloo():
    """host logic of the ray-parallel render (make-it-3d_b200/parallel.py::RayParallel) at world_size 2 on CPU"""
    mp.spawn(_ray_parallel_worker, args=(2, _free_port()), nprocs=2, join=True)
