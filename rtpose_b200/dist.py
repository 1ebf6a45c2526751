"""Data-parallel plumbing for the hot path: frames shard by rank, gradients average through ONE flat fp32 buffer.

Mirrors det3d/core/utils/dist_utils.py:8-57 (`_allreduce_coalesced` / `allreduce_grads`: flatten -> all_reduce ->
div by world size -> copy back) and the DDP wrap at det3d/torchie/apis/train.py:285-291 — minus the reference's
duplicate second all-reduce.  Gradients already live in one flat buffer here, so there is no flatten/copy-back.
The collective is torch.distributed's (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_frames(total_frames, rank, world):
    """Contiguous shard [lo, hi) of `total_frames` for `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(int(total_frames), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_flat(flat, world=None, async_op=False):
    """In-place mean of a flat gradient buffer over all ranks.  Returns the work handle when async_op."""
    if not (dist.is_available() and dist.is_initialized()):
        return None
    world = world or dist.get_world_size()
    if world == 1:
        return None
    if async_op:
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    scale_flat(flat, 1.0 / world)
    return None


def allreduce_grads(params, world=None):
    """Average `.grad` of every parameter over the ranks through one coalesced buffer — the call a trainer makes after
    loss.backward() (det3d/core/utils/dist_utils.py:40-57 `allreduce_grads(..., coalesce=True)`).  No-op when
    torch.distributed is not initialised or the world is one rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = world or dist.get_world_size()
    if world == 1:
        return
    grads = [p.grad for p in params if p.requires_grad and p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    allreduce_flat(flat, world)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def scale_flat(flat, s):
    if flat.is_cuda:
        from . import lib
        lib.call("rtp_scale_f32", flat.data_ptr(), flat.numel(), float(s), torch.cuda.current_stream().cuda_stream)
    else:
        flat.mul_(s)


def broadcast_params(params, src=0):
    """One-time parameter broadcast (DDP does this at wrap time)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for p in params:
            dist.broadcast(p.data if hasattr(p, "data") else p, src)
