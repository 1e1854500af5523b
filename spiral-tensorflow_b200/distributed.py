"""Synchronous data parallelism inside one 8-GPU box (replaces the reference's asynchronous
gRPC parameter-server topology: utils/tf.py:35-50, agent.py:38-59, trainer.py:107-118).

One process per GPU (torchrun), weights replicated, the environment batch sharded as ACTOR
PARTITIONS -- canvases are independent units (each owns its surface, brush and RNG cursor,
envs/mnist.py:89-95), so render / D reward / A2C loss need no data-path collective.  The only
exchange step is one all-reduce of the gradients per optimiser step (NCCL over NVLink/NVSwitch on
GPUs, gloo in the CPU tests):

  * the reference loss is a SUM over the batch (agent.py:182-187), so "average of per-GPU sums"
    is global_sum / world -- the 1/world factor is explicit here
  * clip-by-global-norm (agent.py:231) must see the REDUCED gradient: call it after the all-reduce
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init(backend=None, device=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard(total, rank, world):
    """Actor partition of `total` canvases: rank g owns [start, start+count); counts differ by <= 1."""
    base, rem = divmod(int(total), int(world))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def allreduce_grads_(tensors, world=None, average=True, bucket_bytes=64 << 20):
    """In-place all-reduce (sum, then x 1/world) of a list of gradient tensors, flattened into
    buckets sized for launch latency (NVSwitch makes per-link counts irrelevant)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return tensors
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.mul_(1.0 / world)
        off = 0
        for t in bucket:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
        bucket, size = [], 0

    for t in tensors:
        if bucket and (bucket[0].dtype != t.dtype or size + t.numel() * t.element_size() > bucket_bytes):
            flush()
        bucket.append(t)
        size += t.numel() * t.element_size()
    flush()
    return tensors


def clip_by_global_norm_(tensors, clip_norm):
    """tf.clip_by_global_norm (agent.py:231): g *= clip / max(||g||, clip).  Returns the norm."""
    if not tensors:
        return torch.zeros(())
    norm = torch.sqrt(sum((t.double() ** 2).sum() for t in tensors)).to(tensors[0].dtype)
    scale = clip_norm / torch.clamp(norm, min=clip_norm)
    for t in tensors:
        t.mul_(scale)
    return norm


def allreduce_bn_stats_(sum_and_sumsq, count, world=None):
    """Optional global-batch BatchNorm: all-reduce per-layer (sum x, sum x^2) and the element
    count so that D(x) on a sharded batch equals D(x) on the gathered batch
    (models/discriminator.py:76-79 normalises with the statistics of the batch it is fed)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(sum_and_sumsq, op=dist.ReduceOp.SUM)
        c = torch.tensor([float(count)], dtype=torch.float64, device=sum_and_sumsq.device)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        count = float(c.item())
    mean = sum_and_sumsq[0] / count
    var = sum_and_sumsq[1] / count - mean * mean
    return mean, var.clamp_(min=0), count
