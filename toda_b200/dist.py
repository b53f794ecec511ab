"""Data-parallel plumbing: frame sharding and the one collective of the hot path.

The path shards by frame: rank r voxelizes and runs frames [B*r, B*(r+1)); no activation
ever crosses GPUs.  The only exchange is the average of the backbone's parameter gradients, which the
reference gets from DistributedDataParallel (tools/stage1_cutmix_train.py L142).  Here the 2.7 M floats (~10.8 MB for
VoxelResBackBone8x) live in ONE flat buffer cut into a few buckets in REVERSE parameter order (the backward pass produces
the last layers' gradients first: conv_out + conv4 are 70 % of the bytes); a bucket is packed with one fused copy and
all-reduced (NCCL AVG, asynchronously, on NCCL's own stream) as soon as its last gradient has been produced, so the
collective overlaps the dgrad / wgrad kernels of the earlier layers.  Afterwards every p.grad IS a view of the flat
buffer: no unpack copy, no division kernel.
"""
import os

import torch
import torch.distributed as dist


def shard_frames(global_first_frame, frames_per_rank, rank):
    """First frame index of `rank`'s sub-batch."""
    return global_first_frame + rank * frames_per_rank


class FlatGradBucket:
    """Usage per step:  bucket.zero(); loss.backward(); bucket.all_reduce_mean()."""

    def __init__(self, params, bucket_bytes=3 << 20, overlap=None, group=None):
        if overlap is None:         # TODA_DIST_OVERLAP=0: one all-reduce after the backward pass (A/B measurements)
            overlap = os.environ.get("TODA_DIST_OVERLAP", "1") != "0"
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        p0 = self.params[0]
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        self.views = list(torch.split(self.flat, [p.numel() for p in self.params]))
        self._index = {id(p): i for i, p in enumerate(self.params)}
        # buckets = contiguous parameter ranges [lo, hi), built from the LAST parameter backwards
        self.buckets = []
        hi, size = len(self.params), 0
        for i in range(len(self.params) - 1, -1, -1):
            size += self.params[i].numel() * self.params[i].element_size()
            if size >= bucket_bytes or i == 0:
                self.buckets.append((i, hi))
                hi, size = i, 0
        self._bucket_of = {}
        for b, (lo, hi) in enumerate(self.buckets):
            for i in range(lo, hi):
                self._bucket_of[i] = b
        self._overlap = bool(overlap)
        self._sync = True            # False inside no_sync(): backward passes stay local
        self._hooks = []
        if self._overlap and hasattr(p0, "register_post_accumulate_grad_hook"):
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._reset()

    # ------------------------------------------------------------------------------------------------
    def _world(self):
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    def _reset(self):
        self._pending = [hi - lo for lo, hi in self.buckets]
        self._ready = [False] * len(self.buckets)
        self._next = 0              # buckets are launched strictly in this order on every rank
        self._work = []

    def zero(self):
        for p in self.params:
            p.grad = None
        self._reset()

    def no_sync(self):
        """Context manager (the counterpart of DistributedDataParallel.no_sync): backward passes inside it launch no
        collective -- for a pass that only some ranks run (per-kernel profiling on rank 0) or for gradient accumulation."""
        bucket = self

        class _NoSync:
            def __enter__(self):
                self.prev, bucket._sync = bucket._sync, False

            def __exit__(self, *exc):
                bucket._sync = self.prev
                return False
        return _NoSync()

    def _on_grad(self, p):
        if not self._sync or self._world() == 1:
            return
        i = self._index.get(id(p))
        if i is None:
            return
        b = self._bucket_of[i]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._ready[b] = True
            self._launch_ready()

    def _launch_ready(self):
        while self._next < len(self.buckets) and self._ready[self._next]:
            self._launch(self._next)
            self._next += 1

    def _launch(self, b):
        lo, hi = self.buckets[b]
        src, dst = [], []
        for i in range(lo, hi):
            g = self.params[i].grad
            if g is None:
                self.views[i].zero_()          # a parameter without a gradient on this rank contributes zeros
            else:
                src.append(g.reshape(-1))
                dst.append(self.views[i])
        if dst:
            torch._foreach_copy_(dst, src)     # one fused pack kernel per bucket
        chunk = self.flat[self.views[lo].storage_offset():self.views[hi - 1].storage_offset() + self.views[hi - 1].numel()]
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            self._work.append((dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True), None))
        else:
            self._work.append((dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True), chunk))

    def all_reduce_mean(self, group=None):
        """Finishes the step's collective: launches what the backward hooks could not (buckets holding a parameter that
        got no gradient on this rank), waits, and points every p.grad at its averaged view of the flat buffer (also for
        parameters whose grad was None here: the other ranks' contributions must reach this replica too)."""
        world = self._world()
        if world == 1 or not self._sync:
            return
        for b in range(self._next, len(self.buckets)):
            self._launch(b)
        self._next = len(self.buckets)
        for work, chunk in self._work:
            work.wait()
            if chunk is not None:
                chunk.div_(world)
        self._work = []
        for p, v in zip(self.params, self.views):
            p.grad = v.view_as(p)
