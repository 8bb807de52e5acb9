"""Batch-sharded data parallelism for the inverse-conv path.

Every image's solve is independent and the weight is replicated, so the path shards over the
batch with NO data-path collective; the only exchange is the sum of dW over the shards, i.e.
one all-reduce of the flat gradient bucket per step (NCCL over NVLink on GPUs; the same code
runs over gloo on CPU tensors in tests).  One process per GPU, launched by torchrun -- the
reference uses single-process nn.DataParallel (inf/if_multiGPU_imagenet32.py:410-411).
"""
import ctypes

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


def rank_order_sum(buckets):
    """what the peer all-reduce kernel computes: the buckets summed in rank order 0, 1, ... (float32, left to
    right) -- the same association on every rank, hence bit-identical results everywhere."""
    total = buckets[0].clone()
    for b in buckets[1:]:
        total += b
    return total


class PeerAllReduce:
    """The dW bucket of every rank in symmetric (peer-mapped) memory + the fused all-reduce kernel over it
    (csrc/ifk_comm.cu, include/ifk.h: ifk_allreduce_peer_f32).  torch is used for the plumbing only: the
    symmetric allocation and the rendezvous that exchanges the peer mappings (NVLink / NVSwitch P2P).

        comm = PeerAllReduce(n_floats, device)       # collective: every rank of the default group
        ... kernels write comm.bucket (this rank's partial sums) ...
        comm.allreduce(out)                          # out = sum over ranks, inside the current stream / graph
    """

    def __init__(self, numel, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _native
        self.lib = _native.load()
        self.device = torch.device(device)
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.numel = int(numel)
        pad = (self.numel + 3) // 4 * 4
        flag_words = self.lib.ifk_allreduce_flag_bytes() // 4
        # one symmetric allocation: [bucket | flag block]; rendezvous maps every rank's copy into this process
        self.storage = symm_mem.empty(pad + flag_words, dtype=torch.float32, device=self.device)
        self.storage.zero_()
        self.handle = symm_mem.rendezvous(self.storage, self.group.group_name)
        self.bucket = self.storage[:self.numel]
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == self.storage.data_ptr()
        vp = ctypes.c_void_p
        self._buckets = (vp * self.world)(*[vp(p) for p in ptrs])
        self._flags = (vp * self.world)(*[vp(p + pad * 4) for p in ptrs])
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)                      # every flag block is zero before anyone signals

    def allreduce(self, out, offset=0, numel=None):
        """out[offset : offset + numel] = sum over ranks of bucket[offset : offset + numel] (a slice must start
        at a multiple of 4 floats); enqueued on the current stream, capturable."""
        from . import _native
        numel = self.numel - offset if numel is None else int(numel)
        if offset % 4 or out.dtype != torch.float32 or not out.is_contiguous() or out.numel() < offset + numel:
            raise ValueError("bad slice / output for the peer all-reduce")
        vp = ctypes.c_void_p
        buckets = (vp * self.world)(*[vp(self._buckets[r] + offset * 4) for r in range(self.world)])
        _native.check(self.lib.ifk_allreduce_peer_f32(buckets, self._flags, self.rank, self.world,
                                                      vp(out.data_ptr() + offset * 4), numel,
                                                      _native.current_stream(self.device)))
        return out


class DataParallelStack:
    """An InvConvStack per rank on that rank's shard + the bucket all-reduce after every step."""

    def __init__(self, stack, group=None, average=False):
        self.stack, self.group, self.average = stack, group, average

    def step(self):
        self.stack.step()
        return allreduce_gradients(self.stack.grad_bucket, self.group, self.average)
