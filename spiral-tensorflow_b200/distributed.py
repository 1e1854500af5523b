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


class GradReducer(object):
    """Gradient exchange overlapped with the backward pass, into a buffer the optimiser reads in place.

    All parameters' gradients live in ONE persistent flat buffer (``p.grad`` are views of it, laid out in reverse
    parameter order -- the order in which backward produces them), cut into buckets of ``bucket_bytes``.  A
    post-accumulate-grad hook counts a bucket's parameters down and, when the last one is in, launches that bucket's
    all-reduce (sum) asynchronously on the process group's stream -- NCCL over NVLink/NVSwitch runs under the rest of the
    backward instead of after a ``torch.cat`` -> reduce -> copy-back pass (``allreduce_grads_``).  ``wait()`` joins the
    outstanding collectives (and reduces buckets whose parameters received no gradient); K5 (``FusedTFAdam``) then reads
    the reduced gradients where they are, with the explicit 1/world in its ``grad_scale``.  The hooks issue the same
    sequence of collectives on every rank and every step, so the step can be captured into a CUDA graph."""

    def __init__(self, params, world=None, bucket_bytes=64 << 20):
        self.params = [p for p in params]
        self.world = (dist.get_world_size() if dist.is_initialized() else 1) if world is None else int(world)
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros((total,), dtype=dt, device=dev)
        self.buckets = []                                  # [start, end) element ranges of the flat buffer
        self._bucket_of = {}
        off, b_start, b_elems = 0, 0, 0
        per_bucket = max(1, int(bucket_bytes) // self.flat.element_size())
        self._members = [[]]
        for i in reversed(range(len(self.params))):
            p = self.params[i]
            n = p.numel()
            if b_elems and b_elems + n > per_bucket:
                self.buckets.append((b_start, off))
                self._members.append([])
                b_start, b_elems = off, 0
            p.grad = self.flat[off:off + n].view_as(p)
            self._bucket_of[i] = len(self._members) - 1
            self._members[-1].append(i)
            off += n
            b_elems += n
        self.buckets.append((b_start, off))
        self._pending = [len(m) for m in self._members]
        self._fired = [False] * len(self.buckets)
        self._work = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]

    def _make_hook(self, i):
        def hook(param):
            if param.grad is None or param.grad.data_ptr() != self.flat.data_ptr() + self._offset_bytes(i):
                # something replaced the view (zero_grad(set_to_none=True)): copy into the flat buffer and re-attach
                g = param.grad
                param.grad = self._view(i)
                if g is not None:
                    param.grad.copy_(g)
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._fire(b)
        return hook

    def _view(self, i):
        off = self._offset_bytes(i) // self.flat.element_size()
        return self.flat[off:off + self.params[i].numel()].view_as(self.params[i])

    def _offset_bytes(self, i):
        if not hasattr(self, "_offs"):
            self._offs, off = {}, 0
            for j in reversed(range(len(self.params))):
                self._offs[j] = off * self.flat.element_size()
                off += self.params[j].numel()
        return self._offs[i]

    def _fire(self, b):
        if self._fired[b]:
            return
        self._fired[b] = True
        if self.world > 1:
            s, e = self.buckets[b]
            self._work.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, async_op=True))

    def zero(self):
        """Start of a step: clear the flat buffer (gradients accumulate into it) and re-arm the buckets."""
        self.flat.zero_()
        for i, p in enumerate(self.params):
            if p.grad is None:
                p.grad = self._view(i)
        self._pending = [len(m) for m in self._members]
        self._fired = [False] * len(self.buckets)
        self._work = []

    def wait(self):
        """After backward: reduce the buckets that never filled (parameters without a gradient this step), then make the
        current stream wait for every collective."""
        for b in range(len(self.buckets)):
            self._fire(b)
        for w in self._work:
            w.wait()
        self._work = []

    def collective_bytes(self):
        return self.flat.numel() * self.flat.element_size()


def clip_by_global_norm_(tensors, clip_norm):
    """tf.clip_by_global_norm (agent.py:231): g *= clip / max(||g||, clip).  Returns the norm."""
    if not tensors:
        return torch.zeros(())
    norm = torch.sqrt(sum((t.double() ** 2).sum() for t in tensors)).to(tensors[0].dtype)
    scale = clip_norm / torch.clamp(norm, min=clip_norm)
    for t in tensors:
        t.mul_(scale)
    return norm


def allreduce_bn_sums_(sums, count):
    """The ``global_bn`` hook of ``Discriminator.forward``: sum a layer's per-channel (sum x, sum x^2) [C,2] float64 and the
    element count [1] float64 over the ranks, in place (NCCL on GPUs, gloo in the CPU tests)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(count, op=dist.ReduceOp.SUM)


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
