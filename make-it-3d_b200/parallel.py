"""View-parallel data parallelism for the SDS step: one process per GPU, each rank renders a DISTINCT camera
pose and runs its own SD guidance pass; the only exchange step is one all-reduce (SUM) of the hash-grid + MLP gradients over
NCCL / NVLink before the (identical) optimizer update on every rank.  The reference has no working multi-GPU path
(nerf/utils.py:255-263 is unreachable); the semantics defined here are:

    a G-rank step  ==  single-process accumulation of the same G poses' gradients, then one optimizer step.

The occupancy grid must stay identical on all ranks: `update_extra_state` is run redundantly with a seed shared by all
ranks (same parameters + same jitter -> bit-identical bitfield), so no broadcast is needed.
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def pose_index(step, rank, world_size):
    """Pose consumed by `rank` at global step `step`: consecutive poses are dealt round-robin, so the union over ranks of a
    G-rank run equals the pose sequence of a single-process run (nerf/provider.py:266 index semantics, incl. index % 4 == 0
    being the fixed front view)."""
    return step * world_size + rank


def rank_seed(base_seed, rank):
    """Per-rank RNG stream for view-dependent draws (march jitter, light direction, timestep, SD noises)."""
    return int(base_seed) + 1000003 * int(rank)


def shared_seed(base_seed, iteration):
    """Seed shared by ALL ranks for state that must stay replicated (density-grid jitter)."""
    return (int(base_seed) * 2654435761 + int(iteration)) % (2 ** 62)


class GradientAllReduce:
    """Flat-bucket all-reduce of the two optimizer groups of NeRFNetwork.get_params (nerf/network_tcnn.py:195-206):
    encoder.params (12 196 240 fp32, reduced in place, no copy) and the 6 MLP tensors (6 532 fp32, packed into one bucket).
    Payload 48.8 MB per step -> ~0.1 ms at NVLink-5 line rate."""

    def __init__(self, encoder_params, mlp_params, op="sum"):
        self.encoder_params = encoder_params
        self.mlp_params = list(mlp_params)
        self.op = op
        n = sum(p.numel() for p in self.mlp_params)
        self._bucket = torch.zeros(n, dtype=torch.float32, device=encoder_params.device)

    def __call__(self):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        if self.encoder_params.grad is None:
            self.encoder_params.grad = torch.zeros_like(self.encoder_params)
        work = dist.all_reduce(self.encoder_params.grad, op=dist.ReduceOp.SUM, async_op=True)
        off = 0
        for p in self.mlp_params:
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            self._bucket[off:off + p.numel()].copy_(g.reshape(-1))
            off += p.numel()
        dist.all_reduce(self._bucket, op=dist.ReduceOp.SUM)
        work.wait()
        off = 0
        for p in self.mlp_params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            p.grad.copy_(self._bucket[off:off + p.numel()].view_as(p))
            off += p.numel()
        if self.op == "mean":
            self.encoder_params.grad.div_(world)
            for p in self.mlp_params:
                p.grad.div_(world)


def max_over_ranks(value, device):
    """max of a python float over all ranks (timing rule: report the slowest rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
