"""CUDA-graph capture of the per-micro-batch training work (forward + loss + backward).

A micro-batch issues ~600 small host calls (ctypes launches, autograd bookkeeping, allocator
traffic); on a B200 their GPU work takes ~40 ms per tile while the host needs ~50 ms to issue it.
Capturing the whole forward/backward once and replaying it removes the host from the loop (CUDA
streams and graphs instead of a tracing compiler).  Everything on the path is capturable: the C-ABI
entry points only enqueue work on the current stream, tensor maps travel as kernel parameters, the
cell sort is device-side, no op synchronises.

Weights change between replays (optimizer steps), so the TF32 hi/lo splits of the weights must be
part of the graph: capture runs with the split cache bypassed (``linear.capture_mode``).
Gradients accumulate in place into the parameters' existing ``.grad`` tensors (give them static
storage first, e.g. ``parallel.FlatGradients``); the caller zeroes them between optimizer steps.
"""
import torch

from . import linear as _linear


class GraphedTrainStep:
    def __init__(self, model, loss_fn, *example_inputs, warmup: int = 2):
        """``loss_fn(model, *inputs) -> scalar loss tensor``; ``example_inputs`` fix the shapes."""
        self.model, self.loss_fn = model, loss_fn
        for p in model.parameters():
            if p.requires_grad and p.grad is None:
                p.grad = torch.zeros_like(p)
        self.static_inputs = [torch.empty_like(t) for t in example_inputs]
        for s, t in zip(self.static_inputs, example_inputs):
            s.copy_(t)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):  # lazy handles, kernel attributes, autotuning happen outside the capture
                loss_fn(model, *self.static_inputs).backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._grad_homes = [(p, p.grad.data_ptr()) for p in model.parameters() if p.requires_grad and p.grad is not None]
        self.graph = torch.cuda.CUDAGraph()
        _linear.capture_mode = True
        try:
            with torch.cuda.graph(self.graph):
                loss = loss_fn(model, *self.static_inputs)
                loss.backward()
                self.static_loss = loss.detach()
        finally:
            _linear.capture_mode = False

    def __call__(self, *inputs):
        # the graph accumulates into the gradient tensors that existed at capture time: a parameter whose .grad was
        # replaced since (optimizer.zero_grad(set_to_none=True)) would silently get no update
        for p, ptr in self._grad_homes:
            if p.grad is None or p.grad.data_ptr() != ptr:
                raise RuntimeError("GraphedTrainStep: a parameter's .grad was replaced after capture (use "
                                   "FlatGradients.zero_() / zero_grad(set_to_none=False) between optimizer steps)")
        for s, t in zip(self.static_inputs, inputs):
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_loss
