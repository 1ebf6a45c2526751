"""N > 1 host logic on CPU: two gloo ranks shard frames, average a flat gradient buffer in place and receive the
rank-0 parameters (what bench.py / a DDP-style trainer does around the hot path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rtpose_b200 import dist as rdist
        lo, hi = rdist.shard_frames(33, rank, world)
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)        # "gradients" of this rank
        rdist.allreduce_flat(flat, world)
        params = torch.full((5,), float(rank + 7))
        rdist.broadcast_params([params])
        t3 = torch.ones(3) * rank
        h = rdist.allreduce_flat(t3, world, async_op=True)
        h.wait()
        assert t3.tolist() == [0.5] * 3, t3     # wait() completes the MEAN, not just the sum
        h.wait()
        assert t3.tolist() == [0.5] * 3         # ... exactly once
        ps = [torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(1))]
        ps[0].grad, ps[1].grad = torch.full((2, 3), float(rank)), torch.full((4,), 10.0 * (rank + 1))
        if rank == 1:
            ps[2].grad = torch.ones(1)  # only one rank touched this parameter: the other contributes zeros, same layout
        rdist.allreduce_grads(ps, world)
        assert ps[0].grad.tolist() == [[0.5] * 3] * 2 and ps[1].grad.tolist() == [15.0] * 4 and ps[2].grad.tolist() == [0.5]
        # gradients that are consecutive views of one flat buffer are reduced in place, without flatten / copy-back
        flatg = torch.arange(11, dtype=torch.float32) * (rank + 1)
        qs = [torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(5))]
        qs[0].grad, qs[1].grad = flatg[0:6].view(2, 3), flatg[6:11].view(5)
        assert rdist._common_flat([q.grad for q in qs]) is flatg
        rdist.allreduce_grads(qs, world)
        assert flatg.tolist() == [i * 1.5 for i in range(11)] and qs[1].grad.data_ptr() == flatg[6:].data_ptr()
        # sliced all-reduce (the collective side of the backward overlap): slices are cut from the tail of the buffer at
        # parameter boundaries and reduced as they are reported ready; pre-scaled gradients -> plain sum
        names = [("a", 10), ("b", 30), ("c", 5), ("d", 55), ("e", 20), ("f", 40)]
        fl = torch.arange(160, dtype=torch.float32) * (rank + 1) / world
        sar = rdist.SlicedAllReduce(fl, names, 3, world)
        assert sar.spans == [(120, 160), (100, 120), (0, 100)] and sar.groups == [["f"], ["e"], ["a", "b", "c", "d"]]
        for k in range(len(sar.spans)):
            sar.on_ready(k)
        sar.join()
        assert fl.tolist() == [i * 1.5 for i in range(160)]
        out.put((rank, lo, hi, flat.tolist(), params.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allreduce_and_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, f0, p0), (r1, lo1, hi1, f1, p1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 17, 17, 33)
    want = [i * 1.5 for i in range(10)]  # mean of 1x and 2x
    assert f0 == want and f1 == want
    assert p0 == [7.0] * 5 and p1 == [7.0] * 5


def test_single_process_is_a_noop():
    from rtpose_b200 import dist as rdist
    t = torch.ones(4)
    assert rdist.allreduce_flat(t) is None and t.tolist() == [1.0] * 4
