"""Batch-sharded data parallelism for the inverse-conv path.

Every image's solve is independent and the weight is replicated, so the path shards over the
batch with NO data-path collective; the only exchange is the sum of dW over the shards, i.e.
one all-reduce of the flat gradient bucket per step (NCCL over NVLink on GPUs; the same code
runs over gloo on CPU tensors in tests).  One process per GPU, launched by torchrun -- the
reference uses single-process nn.DataParallel (inf/if_multiGPU_imagenet32.py:410-411).
"""
import torch
import torch.distributed as dist


def shard_bounds(global_batch, world_size, rank):
    """Contiguous shard [start, stop) of a global batch; sizes differ by at most one and the
    first `global_batch % world_size` ranks take the larger shards (ragged batches allowed)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank %d / world size %d" % (rank, world_size))
    base, extra = divmod(int(global_batch), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_gradients(bucket, group=None, average=False):
    """Sum (or average) the flat dW bucket over the data-parallel group, in place."""
    if not dist.is_available() or not dist.is_initialized():
        return bucket
    world = dist.get_world_size(group)
    if world == 1:
        return bucket
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    if average:
        bucket.div_(world)
    return bucket


def max_over_ranks(value, device=None, group=None):
    """max of a python float over the group (timing: a step is as slow as its slowest rank)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class DataParallelStack:
    """An InvConvStack per rank on that rank's shard + the bucket all-reduce after every step."""

    def __init__(self, stack, group=None, average=False):
        self.stack, self.group, self.average = stack, group, average

    def step(self):
        self.stack.step()
        return allreduce_gradients(self.stack.grad_bucket, self.group, self.average)
