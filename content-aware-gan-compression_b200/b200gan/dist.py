"""Data-parallel plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink on the box,
gloo in the CPU tests).  Mirrors the helper names of the reference's Miscellaneous/distributed.py
(:9-126) and makes `gather_grad` real: the reference all-reduces ~90 parameter tensors one by one
(:57-66) and never calls it; here the student's gradients live in ONE flat fp32 bucket that is
all-reduced once per step and consumed in place by the fused Adam kernel.
"""
from __future__ import annotations

import os
import pickle
from typing import Iterable, List, Sequence

import torch
from torch import distributed as dist


def _active() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_rank() -> int:
    return dist.get_rank() if _active() else 0


def get_world_size() -> int:
    return dist.get_world_size() if _active() else 1


def synchronize():
    if _active() and dist.get_world_size() > 1:
        dist.barrier()


def init_from_env(backend: str | None = None) -> int:
    """Initialise the default process group from torchrun's RANK / WORLD_SIZE / MASTER_* variables.
    Returns the local rank.  No-op for single-process runs."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not _active():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        kw = {'device_id': torch.device('cuda', local)} if backend == 'nccl' else {}
        dist.init_process_group(backend=backend, **kw)
    return local


def reduce_sum(tensor: torch.Tensor) -> torch.Tensor:
    if not _active():
        return tensor
    tensor = tensor.clone()
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def reduce_loss_dict(loss_dict):
    if get_world_size() < 2:
        return loss_dict
    with torch.no_grad():
        keys = sorted(loss_dict.keys())
        losses = torch.stack([loss_dict[k] for k in keys], 0)
        dist.reduce(losses, dst=0)
        if dist.get_rank() == 0:
            losses /= get_world_size()
        return {k: v for k, v in zip(keys, losses)}


def all_gather(data):
    """Gather arbitrary picklable objects from every rank (reference :69-101)."""
    if get_world_size() == 1:
        return [data]
    out = [None] * get_world_size()
    dist.all_gather_object(out, data)
    return out


class FlatBucket:
    """All parameters of a module re-homed into one flat fp32 buffer, gradients into a second one.

    `param.data` and `param.grad` become views, so autograd accumulates straight into the bucket and
    one collective + one optimizer kernel cover the whole model.  The student generator is 5.57 M
    parameters = 22.3 MB: a single NVLink all-reduce instead of ~90 small ones.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], with_grads: bool = True, all_params: bool = False):
        """with_grads=False: parameter bucket only (the EMA copy of the generator, train.py:124-129).
        all_params=True: take every parameter, trainable or not (an EMA model is usually frozen)."""
        self.params: List[torch.nn.Parameter] = [p for p in params if (all_params or p.requires_grad)]
        assert self.params, 'no trainable parameters'
        dev, dt = self.params[0].device, self.params[0].dtype
        # 4-float alignment of every segment keeps 128-bit accesses legal for any consumer of a view
        self.offsets = []
        n = 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.flat_param = torch.zeros(n, device=dev, dtype=dt)
        self.flat_grad = torch.zeros(n, device=dev, dtype=dt) if with_grads else None
        for p, off in zip(self.params, self.offsets):
            seg = self.flat_param[off:off + p.numel()].view_as(p)
            seg.copy_(p.data)
            p.data = seg
            if with_grads:
                p.grad = self.flat_grad[off:off + p.numel()].view_as(p)

    def zero_grad(self):
        self.flat_grad.zero_()
        for p, off in zip(self.params, self.offsets):   # re-attach in case something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + off * 4:
                p.grad = self.flat_grad[off:off + p.numel()].view_as(p)

    # -- "detached" protocol: one multi-tensor copy per step instead of one accumulation kernel per parameter.
    # With .grad pre-attached to the bucket autograd ADDS every gradient into its view (one tiny launch per
    # parameter, ~135 per student step, after a 22 MB memset).  detach_grads() sets .grad = None, so backward
    # hands each gradient tensor over without a kernel; pack_grads() then moves all of them into the bucket
    # with a single batched copy and re-attaches the views.
    def detach_grads(self):
        for p in self.params:
            p.grad = None

    def pack_grads(self):
        views = [self.flat_grad[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]
        have = [(v, p.grad) for v, p in zip(views, self.params)
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        if any(p.grad is None for p in self.params):
            self.flat_grad.zero_()           # parameters the loss did not reach keep a zero gradient
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for p, v in zip(self.params, views):
            p.grad = v

    def allreduce_mean_(self, async_op: bool = False):
        """Sum the bucket over ranks (the division by world size is folded into the optimizer
        kernel's grad_scale).  Semantics of gather_grad (reference :57-66)."""
        if get_world_size() == 1:
            return None
        return dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, async_op=async_op)


def gather_grad(params: Sequence[torch.nn.Parameter]):
    """Drop-in for the reference's gather_grad: average .grad over ranks, bucketed into one message."""
    world = get_world_size()
    if world == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def shard_batches(n_batches: int, rank: int | None = None, world: int | None = None) -> List[int]:
    """Saliency pass sharding: whole reference batches stay together (the metric takes
    |.| after the in-batch sum, Util/content_aware_pruning.py:184-196); rank r owns batches r, r+N, ..."""
    rank = get_rank() if rank is None else rank
    world = get_world_size() if world is None else world
    return list(range(rank, n_batches, world))


def gather_scores_in_batch_order(local: dict, n_batches: int):
    """local: {batch index: [np.ndarray per layer]}.  Returns on every rank the list over batches in
    batch order, so that the final sum over batches (prune.py:45-46) is evaluated in the same order
    as the single-GPU run and stays bit-identical to it."""
    merged = {}
    for part in all_gather(local):
        merged.update(part)
    assert sorted(merged) == list(range(n_batches)), 'missing saliency batches'
    return [merged[i] for i in range(n_batches)]
