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
