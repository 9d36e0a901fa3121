"""Data-parallel trainer for the B200 path -- the multi-GPU, batched counterpart of the
reference's ``Trainer`` (trainer.py:8-89) and its driver loop (train.py:148-153).

The reference trains at batch 1 (tiles have different point counts, conf/model/tomosar2height.yaml:40) and
accumulates un-normalised tile gradients over ``optimize_every`` tiles before one optimizer step
(trainer.py:70-89).  Here the same accumulation window is processed in micro-batches of several tiles (dense
``(B, N, 3)`` batches replay a captured CUDA graph; ragged batches -- ``RaggedCloud`` or a list of ``(N_i, 3)``
tiles -- run eagerly), every rank works on its own tiles, and ONE flat fp32 all-reduce (SUM, like the
reference's un-normalised accumulation) precedes the optimizer step.  Loss per tile = trainer.py:63-69:
L1 on heights (+ weight_ce * BCE-with-logits on the footprint head).
"""
import torch
import torch.nn.functional as F

from .parallel import FlatGradients
from .topology import RaggedCloud


class Trainer:
    def __init__(self, model, optimizer, micro_batch=4, use_cuda_graph=True, use_footprint=False, weight_ce=10.0,
                 optimize_every=64):
        self.model, self.optimizer = model, optimizer
        self.micro_batch, self.use_cuda_graph = micro_batch, use_cuda_graph
        self.use_footprint, self.weight_ce = use_footprint, weight_ce
        self.optimize_every = optimize_every
        self.flat = FlatGradients(model)
        self._graphs = {}          # input shapes -> GraphedTrainStep
        self._accumulated = 0      # tiles since the last optimizer step (train_step API)
        self._ar_events = None

    # -- loss of a micro-batch: sum over its tiles of the per-tile loss (trainer.py:63-69) -----------------
    def micro_loss(self, model, cloud, dsm, image=None):
        pa, pb = model(input_cloud=cloud, input_image=image)
        loss = (pa.squeeze(-1) - dsm).abs().mean(dim=(1, 2)).sum()
        if self.use_footprint:
            target = (dsm > 0.0001).to(pa.dtype)
            bce = F.binary_cross_entropy_with_logits(pb.squeeze(-1), target, reduction="none").mean(dim=(1, 2)).sum()
            loss = loss + self.weight_ce * bce
        return loss

    def _graph_for(self, inputs):
        from .graph import GraphedTrainStep
        key = tuple((tuple(t.shape), t.dtype) for t in inputs)
        if key not in self._graphs:
            self._graphs[key] = GraphedTrainStep(self.model, self.micro_loss, *inputs)
            self.flat.verify()
        return self._graphs[key]

    def _backward_micro(self, inputs, eager):
        dense = all(isinstance(t, torch.Tensor) for t in inputs)
        if self.use_cuda_graph and dense and not eager:
            return self._graph_for(inputs)(*inputs)
        loss = self.micro_loss(self.model, *inputs)
        loss.backward()
        return loss.detach()

    def _reduce_and_step(self):
        """accumulation-aware: the collective runs once, right before the optimizer step"""
        if self.flat.flat.is_cuda:  # device time of the collective, for the bench line
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.flat.all_reduce()
            e1.record()
            self._ar_events = (e0, e1)
        else:
            self.flat.all_reduce()
        self.optimizer.step()
        self._accumulated = 0

    def last_allreduce_ms(self):
        if self._ar_events is None:
            return None
        torch.cuda.synchronize()
        return self._ar_events[0].elapsed_time(self._ar_events[1])

    # -- one optimizer step over a whole batch of this rank's tiles ---------------------------------------
    def train_batch(self, cloud, dsm, image=None, eager=False):
        """cloud (T, N, 3) dense, or a list of (N_i, 3) tiles (ragged); dsm (T, S, S); image (T, 3, S, S) or None.
        Forward + loss + backward in micro-batches, ONE gradient all-reduce, optimizer step.  Returns the summed loss."""
        self.flat.zero_()
        ragged = isinstance(cloud, (list, tuple))
        T = len(cloud) if ragged else cloud.shape[0]
        total = torch.zeros((), device=dsm.device)
        for i in range(0, T, self.micro_batch):
            j = min(i + self.micro_batch, T)
            c = RaggedCloud.from_list(list(cloud[i:j])) if ragged else cloud[i:j]
            inputs = (c, dsm[i:j]) + (() if image is None else (image[i:j],))
            total += self._backward_micro(inputs, eager)
        self._reduce_and_step()
        return total

    # -- the reference's call: one tile (or small batch) per call, step every `optimize_every` tiles ------
    def train_step(self, data):
        """trainer.py:47-89: ``data`` = {'inputs': (B, N, 3) | RaggedCloud | list of tiles, 'dsm': (B, S, S),
        'image': optional}.  Gradients accumulate across calls; after ``optimize_every`` tiles (counted over this
        rank) the gradients of all ranks are summed and the optimizer steps.  Returns (loss, stepped)."""
        cloud, dsm, image = data.get("inputs"), data.get("dsm"), data.get("image")
        if self._accumulated == 0:
            self.flat.zero_()
        if isinstance(cloud, (list, tuple)):
            cloud = RaggedCloud.from_list(list(cloud))
        n_tiles = cloud.n_tiles if isinstance(cloud, RaggedCloud) else cloud.shape[0]
        inputs = (cloud, dsm) + (() if image is None else (image,))
        loss = self._backward_micro(inputs, eager=isinstance(cloud, RaggedCloud))
        self._accumulated += n_tiles
        stepped = self._accumulated >= self.optimize_every
        if stepped:
            self._reduce_and_step()
        return loss, stepped
