"""Data-parallel plumbing: frame sharding and the one collective of the hot path.

The path shards by frame: rank r voxelizes and runs frames [B*r, B*(r+1)); no activation
ever crosses GPUs.  The only exchange is the average of the backbone's parameter gradients, which the
reference gets from DistributedDataParallel (tools/stage1_cutmix_train.py L142).  Here it is one flat bucket
(2.7 M floats ~ 10.8 MB for VoxelResBackBone8x) all-reduced once per step over NCCL / NVLink.
"""
import torch
import torch.distributed as dist


def shard_frames(global_first_frame, frames_per_rank, rank):
    """First frame index of `rank`'s sub-batch."""
    return global_first_frame + rank * frames_per_rank


class FlatGradBucket:
    """Views every parameter's .grad into one contiguous buffer so that the step's collective is a single
    all-reduce with no flatten / unflatten copies."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(dist.get_world_size(group))
