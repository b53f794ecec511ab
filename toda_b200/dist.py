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
    """One flat buffer for the step's only collective.  Gradients are produced by autograd into fresh tensors
    (p.grad is reset to None every step, so no accumulate kernels run); `all_reduce_mean` packs them with one
    concatenation, all-reduces the flat buffer once over NCCL and scatters the averages back in place."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        self.views = list(torch.split(self.flat, [p.numel() for p in self.params]))

    def zero(self):
        for p in self.params:
            p.grad = None

    def all_reduce_mean(self, group=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch.cat([g.reshape(-1) for g in grads], out=self.flat)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(dist.get_world_size(group))
        torch._foreach_copy_([g.view(-1) for g in grads], self.views)
