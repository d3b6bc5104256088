"""Whole-step CUDA graphs for the launch-bound training loop.

The reference's inner loop (``topological_training/train.py:107-116``: ``zero_grad`` -> forward ->
SmoothL1 -> ``backward`` -> ``step``) is ~60 small kernels at batch 512-1024; on a B200 each is a
few microseconds, so the step is bound by Python / launch latency, not by the GPU.  Every entry
point of libqot_b200 enqueues on the caller's stream without synchronising, so the whole step --
CSR build, both conv layers, pooling head, loss, all backward kernels (graph A) and the SGD update
(graph B) -- is captured once and replayed per batch, with the flat NCCL gradient all-reduce issued
eagerly between the two replays.

Static shapes only (N, E, B fixed, e.g. batches of one topology such as BASELINE cfg 1/3); a batch
of another shape raises.  Dropout masks drawn with ``torch.rand`` inside the captured region advance
correctly on replay (torch's CUDA generator registers with the graph).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .batch import Batch, _FIELDS


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, criterion: Callable,
                 example: Batch, target_of: Optional[Callable] = None, ddp=None, warmup: int = 3):
        """``model(batch)`` -> prediction; ``criterion(pred, target_of(batch))`` -> scalar loss.
        ``ddp``: a :class:`~.distributed.GraphDataParallel` wrapping ``model`` (its zero / flat
        all-reduce are captured too); ``None`` for single-GPU training."""
        if not example.edge_index.is_cuda:
            raise RuntimeError("GraphedTrainStep needs a CUDA batch (no CPU path)")
        self.model, self.opt, self.crit, self.ddp = model, optimizer, criterion, ddp
        self.target_of = target_of or (lambda b: b.y.view(-1, 3))
        self.static = Batch(num_graphs=example.num_graphs, lut_col=example.lut_col,
                            **{k: (getattr(example, k).clone() if getattr(example, k) is not None else None)
                               for k in _FIELDS})
        self.shapes = {k: tuple(getattr(example, k).shape) for k in _FIELDS if getattr(example, k) is not None}
        # largest graph of the batch: sizes the shared memory of the block-per-graph kernels; read once here
        # (one host sync for a foreign batch), a constant of the captured step afterwards
        from . import ops
        self.static.max_nodes, self.static.max_edges = ops.batch_max_sizes(example)
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for _ in range(max(warmup, 1)):          # allocates workspaces / optimizer state eagerly
                self._front()
                self._exchange()
                self.opt.step()
            torch.cuda.synchronize()
            # graph A: zero -> forward -> loss -> backward (-> gather into the flat buffer);
            # graph B: optimizer update.  The gradient all-reduce runs between the two, eagerly on
            # the same stream: a collective inside a captured region would tie every rank's capture
            # and replay to NCCL's internal streams for no gain (it is one 21 KB call per step).
            self.graph_a = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_a, stream=self.stream):
                self.loss = self._front()
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b, stream=self.stream):
                self.opt.step()
        torch.cuda.current_stream().wait_stream(self.stream)

    def _front(self) -> torch.Tensor:
        self.static._cache = {}                      # the CSR of the batch is rebuilt inside the step
        (self.ddp or self.opt).zero_grad(set_to_none=True)   # backward writes fresh gradients
        loss = self.crit((self.ddp or self.model)(self.static), self.target_of(self.static))
        loss.backward()
        if self.ddp is not None:
            self.ddp.grads.gather()                  # one concatenation kernel; p.grad -> flat views
        return loss.detach()

    def _exchange(self) -> None:
        if self.ddp is not None:
            self.ddp.grads.all_reduce_mean()         # eager NCCL all-reduce of the flat buffer

    def step(self, batch: Batch) -> torch.Tensor:
        """Copies ``batch`` into the static buffers and replays the captured step; returns the
        (device) loss of this step -- reading it on the host is the caller's only sync."""
        for k, shape in self.shapes.items():
            t = getattr(batch, k)
            if t is None or tuple(t.shape) != shape:
                raise RuntimeError(f"GraphedTrainStep was captured for {k}{shape}; got "
                                   f"{None if t is None else tuple(t.shape)} (static shapes only)")
        mn, me = getattr(batch, "max_nodes", None), getattr(batch, "max_edges", None)
        if (mn is not None and mn > self.static.max_nodes) or (me is not None and me > self.static.max_edges):
            raise RuntimeError(f"GraphedTrainStep was captured for graphs of <= {self.static.max_nodes} nodes / "
                               f"{self.static.max_edges} edges; got {mn} / {me}")
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for k in self.shapes:
                getattr(self.static, k).copy_(getattr(batch, k), non_blocking=True)
            self.graph_a.replay()
            self._exchange()
            self.graph_b.replay()
        torch.cuda.current_stream().wait_stream(self.stream)
        return self.loss
