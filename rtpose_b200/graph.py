"""CUDA-graph capture of a whole training / inference step.

One HRRadarPose training step is ~930 kernel launches through the C ABI; enqueueing them from Python costs about as
much host time as the B200 needs to execute them.  Everything in the step is static from one iteration to the next
(shapes, buffer addresses — the engine's pool recycles the same buffers — and the launch sequence), so the step is
captured once and replayed; per-iteration scalars (the one-cycle lr / momentum) live in device memory
(optim.FlatAdam.set_hyper).  The side stream used for weight gradients (ops.conv_wgrad_async) forks from and joins
the capturing stream, so it is captured as a parallel branch of the graph.
"""
import torch


class StepGraph:
    """g = StepGraph(fn, warmup=3); g() replays.  `fn` must read its inputs from fixed device buffers and must not
    synchronise or touch host memory.  Its return value (device tensors) is kept and returned by every replay."""

    def __init__(self, fn, warmup=3, high_priority=True):
        self.fn = fn
        self.high_priority = high_priority
        self.graph = None
        self.result = None
        self.warmup = warmup

    def capture(self):
        from . import ops
        dev = torch.cuda.current_device()
        s = ops.named_stream(dev, "graph_warmup")
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):  # warm-up off the default stream: allocations and lazy initialisation happen here
            for _ in range(self.warmup):
                self.result = self.fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # capture on a high-priority stream: kernel nodes inherit it, so when a main-chain kernel and a weight gradient
        # of the (default-priority) side stream are both ready, the main chain's blocks are dispatched first
        cs = ops.named_stream(dev, "graph_capture", priority=-1) if self.high_priority else ops.named_stream(dev, "graph_capture_lo")
        with torch.cuda.graph(g, stream=cs):
            self.result = self.fn()
        self.graph = g
        return self

    def __call__(self):
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.result
