"""CUDA-graph capture of the per-micro-batch training work (forward + loss + backward).

A micro-batch issues ~600 small host calls (ctypes launches, autograd bookkeeping, allocator
traffic); on a B200 their GPU work takes ~40 ms per tile while the host needs ~50 ms to issue it.
Capturing the whole forward/backward once and replaying it removes the host from the loop (CUDA
streams and graphs instead of a tracing compiler).  Everything on the path is capturable: the C-ABI
entry points only enqueue work on the current stream, tensor maps travel as kernel parameters, the
cell sort is device-side, no op synchronises.

Weights change between replays (optimizer steps) but only once per optimizer step, not once per micro-batch: the
hi/lo operand splits of the PARAMETERS are hoisted out of the graph into persistent buffers, which are rewritten
(eagerly, before the next replay) whenever a parameter's version counter has moved; splits of temporaries (padded or
sliced weights) are recomputed inside the graph (``linear.capture_mode``).  A write that does not bump the version
(``p.data.copy_``) needs an explicit ``refresh_weights()``.
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
        self._params = [p for p in model.parameters()]
        self._recipes = {}       # hoisted parameter splits: key -> (param, builder, f16 flavour, persistent buffers)
        self._weights_token = None
        _linear.static_params = {p.data_ptr(): p for p in self._params}
        _linear.static_recipes = self._recipes
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                # lazy handles, kernel attributes, autotuning happen outside the capture; the persistent buffers of
                # the hoisted weight splits are allocated here, from the ordinary memory pool
                for _ in range(max(warmup, 1)):
                    loss_fn(model, *self.static_inputs).backward()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._grad_homes = [(p, p.grad.data_ptr()) for p in model.parameters() if p.requires_grad and p.grad is not None]
            self.graph = torch.cuda.CUDAGraph()
            _linear.capture_mode = True
            with torch.cuda.graph(self.graph):
                loss = loss_fn(model, *self.static_inputs)
                loss.backward()
                self.static_loss = loss.detach()
        finally:
            _linear.capture_mode = False
            _linear.static_params = _linear.static_recipes = None

    def refresh_weights(self, force=True):
        """bring the hoisted operand splits up to date with the parameters (automatic when a version counter moved)"""
        token = tuple(p._version for p in self._params)
        if force or token != self._weights_token:
            _linear.refresh_static_splits(self._recipes)
            self._weights_token = token

    def __call__(self, *inputs):
        # the graph accumulates into the gradient tensors that existed at capture time: a parameter whose .grad was
        # replaced since (optimizer.zero_grad(set_to_none=True)) would silently get no update
        for p, ptr in self._grad_homes:
            if p.grad is None or p.grad.data_ptr() != ptr:
                raise RuntimeError("GraphedTrainStep: a parameter's .grad was replaced after capture (use "
                                   "FlatGradients.zero_() / zero_grad(set_to_none=False) between optimizer steps)")
        self.refresh_weights(force=False)
        for s, t in zip(self.static_inputs, inputs):
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_loss
