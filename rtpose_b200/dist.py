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


class _MeanWork:
    """Handle of an asynchronous mean all-reduce: wait() also applies the 1/world scale (once)."""

    def __init__(self, work, flat, world):
        self.work, self.flat, self.world = work, flat, world

    def wait(self):
        if self.work is not None:
            self.work.wait()
            scale_flat(self.flat, 1.0 / self.world)
            self.work = None
        return True


def allreduce_flat(flat, world=None, async_op=False, prescaled=False):
    """In-place mean of a flat gradient buffer over all ranks.  async_op: returns a handle whose wait() completes the
    mean (sum, then 1/world).  prescaled: the caller already folded 1/world in (e.g. into the loss gradient seed), so
    only the sum is taken."""
    if not (dist.is_available() and dist.is_initialized()):
        return None
    world = world or dist.get_world_size()
    if world == 1:
        return None
    if async_op:
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
        return work if prescaled else _MeanWork(work, flat, world)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if not prescaled:
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
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    # every rank must reduce the same layout: a parameter without a gradient on this rank contributes zeros
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
    base = _common_flat(grads)
    if base is not None:  # the gradients are already views of one flat buffer (det3d_compat hands them out that way)
        allreduce_flat(base, world)
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    allreduce_flat(flat, world)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


class SlicedAllReduce:
    """All-reduce of a flat gradient buffer in a few contiguous slices, each launched on a communication stream as soon
    as the backward pass has issued the last gradient of the slice (engine.Engine.set_grad_groups), so the collective
    runs under the rest of backward instead of after it.  SURVEY.md §8e; replaces DDP's bucketed overlap
    (det3d/torchie/apis/train.py:285-291) for the flat-buffer trainer.

    `names_in_flat_order`: [(name, numel)] in the order the parameters sit in `flat` — state_dict order, i.e. roughly
    forward order, so backward completes the TAIL of the buffer first and the slices are cut from the back.
    The mean is obtained by seeding the backward pass with 1/world (Engine.loss(grad_scale=1/world)): the collective is a
    plain sum and no scale pass follows.  Everything here is stream-ordered and capturable in a CUDA graph."""

    def __init__(self, flat, names_in_flat_order, nslices=3, world=None, extra_streams=(), fractions=None):
        self.flat = flat
        self.world = world or (dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1)
        total = sum(n for _, n in names_in_flat_order)
        assert total == flat.numel()
        # cut points snapped to parameter boundaries; slice 0 = tail of the buffer.  `fractions` (of the element count, in
        # launch order) default to equal parts; a small LAST slice keeps the collective that cannot overlap short.
        fr = list(fractions) if fractions is not None else [1.0 / nslices] * nslices
        nslices = len(fr)
        marks, c = [], 0.0
        for f in fr[::-1][:-1]:
            c += f
            marks.append(c * total)
        bounds, acc = [0], 0
        for name, n in names_in_flat_order:
            acc += n
            if len(bounds) <= len(marks) and acc >= marks[len(bounds) - 1]:
                bounds.append(acc)
        if bounds[-1] != total:
            bounds.append(total)
        spans = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)][::-1]
        self.spans = spans
        self.groups = []
        for lo, hi in spans:
            o, g = 0, []
            for name, n in names_in_flat_order:
                if lo <= o < hi:
                    g.append(name)
                o += n
            self.groups.append(g)
        if flat.is_cuda:
            from . import ops
            self.stream = ops.named_stream(flat.device, "comm")
        else:
            self.stream = None
        self.extra_streams = list(extra_streams)
        self.launched = []

    def attach(self, engine):
        self.engine = engine
        engine.set_grad_groups(self.groups, self.on_ready)
        return self

    def on_ready(self, k):
        """Called by Engine.backward when every gradient of slice k has been issued."""
        lo, hi = self.spans[k]
        self.launched.append(k)
        if self.world == 1 or not (dist.is_available() and dist.is_initialized()):
            return
        piece = self.flat[lo:hi]
        if self.stream is None:
            dist.all_reduce(piece, op=dist.ReduceOp.SUM)
            return
        from . import ops
        cur = torch.cuda.current_stream(self.flat.device)
        self.stream.wait_stream(cur)
        for st in ops.wgrad_streams(self.flat.device):   # weight gradients (and their split-K reductions) queued so far
            self.stream.wait_stream(st)
        eng = getattr(self, "engine", None)
        for st in list(self.extra_streams) + (list(eng._bstreams.values()) if eng is not None else []):
            self.stream.wait_stream(st)             # branch streams (GroupNorm parameter gradients of side branches)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(piece, op=dist.ReduceOp.SUM)

    def join(self):
        """The current stream waits for every slice's collective (call before the optimizer step)."""
        self.launched = []
        if self.stream is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)


def _common_flat(grads):
    """The flat tensor the gradients are consecutive, gap-free views of, or None."""
    b = grads[0]._base if grads[0]._base is not None else None
    if b is None or b.dim() != 1 or not b.is_contiguous():
        return None
    o = 0
    for g in grads:
        if g._base is not b or not g.is_contiguous() or g.storage_offset() != b.storage_offset() + o:
            return None
        o += g.numel()
    return b if o == b.numel() else None


def scale_flat(flat, s):
    if flat.is_cuda:
        from . import lib
        lib.call("rtp_scale_f32", flat.data_ptr(), flat.numel(), float(s), torch.cuda.current_stream().cuda_stream)
    else:
        flat.mul_(s)


def broadcast_params(params, src=0):
    """One-time parameter broadcast (DDP does this at wrap time)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        with torch.no_grad():  # broadcast INTO the tensor (not .data) so its version counter moves and weight packs rebuild
            for p in params:
                dist.broadcast(p, src)
