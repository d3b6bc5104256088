"""Whole-step CUDA graphs for the launch-bound training loop.

The reference's inner loop (``topological_training/train.py:107-116``: ``zero_grad`` -> forward ->
SmoothL1 -> ``backward`` -> ``step``) is ~60 small kernels at batch 512-1024; on a B200 each is a
few microseconds, so the step is bound by Python / launch latency, not by the GPU.  Every entry
point of libqot_b200 enqueues on the caller's stream without synchronising, so the whole step --
CSR build, both conv layers, pooling head, loss, all backward kernels, the gradient exchange and the
optimizer update -- is captured once and replayed per batch.

Gradient exchange under data parallelism (``ddp=GraphDataParallel(model)``), ``exchange=``:

* ``"peer"`` (default for a plain ``torch.optim.SGD``): the exchange AND the optimizer update are ONE
  native launch inside the step graph -- a one-shot all-reduce over NVLink peer memory fused with the
  SGD rule, hyper-parameters read from device memory (``distributed.FusedSGDStep``, csrc/ddp_step.cu);
  also used single-GPU (no peers), where it replaces the optimizer's three foreach launches;
* ``"graph"``: the flat NCCL all-reduce is captured INSIDE the step graph -- one replay per
  step, no eager launch between kernels;
* ``"eager"``: graph A (zero -> forward -> loss -> backward -> gather) / eager all-reduce / graph B
  (optimizer), the round-1 arrangement, kept for NCCL builds that refuse capture.

Constructing the object does NOT move the model: parameters, buffers and optimizer state are
snapshotted before the warm-up steps that capture needs and restored in place afterwards.  The
optimizer's hyper-parameters (lr after ``StepLR.step()``, momentum, weight decay ...) are baked into
a captured update as host scalars, so ``step()`` compares them with the values captured and
re-captures the update when a scheduler changed them.

Static shapes only (N, E, B fixed, e.g. batches of one topology such as BASELINE cfg 1/3); a batch
of another shape raises (ragged batches: :class:`GraphedStepCache`, one captured graph per shape).  The step is
captured on its own stream: do not keep the (non-detached) loss of an earlier EAGER step alive while constructing
it -- a live autograd graph pins the parameters' AccumulateGrad nodes to the stream of that eager step, and the
capture then fails with cudaErrorStreamCaptureImplicit.  Dropout masks drawn with ``torch.rand`` inside the captured region advance
correctly on replay (torch's CUDA generator registers with the graph).
"""
from __future__ import annotations

import copy
import os
from typing import Callable, Optional

import torch

from .batch import Batch, _FIELDS

_HYPER_KEYS = ("lr", "momentum", "dampening", "weight_decay", "nesterov", "maximize", "betas", "eps", "amsgrad")


def _hyper(opt: torch.optim.Optimizer):
    return tuple(tuple((k, float(g[k]) if isinstance(g[k], (int, float)) else repr(g[k]))
                       for k in _HYPER_KEYS if k in g) for g in opt.param_groups)


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, criterion: Callable,
                 example: Batch, target_of: Optional[Callable] = None, ddp=None, warmup: int = 3,
                 exchange: Optional[str] = None, loss_of: Optional[Callable] = None, shared=None, pool=None,
                 borrow_inputs: bool = False):
        """``model(batch)`` -> prediction; ``criterion(pred, target_of(batch))`` -> scalar loss -- or, for models whose
        output needs more than that (LightpathGNN returns ``(out, lut_batch)``), ``loss_of(model_or_ddp, batch)`` ->
        scalar loss.  ``ddp``: a :class:`~.distributed.GraphDataParallel` wrapping ``model``; ``None`` for
        single-GPU training.  ``shared`` / ``pool``: set by :class:`GraphedStepCache` (one flat gradient buffer,
        one fused optimizer tail and one graph memory pool for all its captured shapes).  ``borrow_inputs``: capture
        the step on the example batch's OWN tensors instead of private copies -- ``step`` then accepts only that very
        batch (same storage) and copies nothing: for datasets resident in HBM whose batches come back unchanged."""
        if not example.edge_index.is_cuda:
            raise RuntimeError("GraphedTrainStep needs a CUDA batch (no CPU path)")
        self.model, self.opt, self.crit, self.ddp = model, optimizer, criterion, ddp
        self.target_of = target_of or (lambda b: b.y.view(-1, 3))
        self.loss_of = loss_of
        self.pool = pool
        from .distributed import FlatGradBuffer, FusedSGDStep, world_info
        exchange = exchange or os.environ.get("QOT_DDP_EXCHANGE", "peer")
        if exchange not in ("peer", "graph", "eager"):
            raise ValueError(f"exchange must be 'peer', 'graph' or 'eager', got {exchange!r}")
        multi = ddp is not None and world_info()[1] > 1
        self.fused = None
        if shared is not None:
            self._grads, self.fused = shared
            exchange = "peer" if self.fused is not None else exchange
        elif exchange == "peer":
            if FusedSGDStep.supports(optimizer):
                try:
                    self._grads = ddp.grads if ddp is not None else FlatGradBuffer(model.parameters())
                    self.fused = FusedSGDStep(optimizer, self._grads, distributed=ddp is not None)
                except Exception as e:                # noqa: BLE001 -- e.g. no peer mapping between the ranks
                    if multi:
                        import warnings
                        warnings.warn(f"GraphedTrainStep: peer exchange unavailable ({e}); using the captured NCCL all-reduce")
                    self.fused = None
            if self.fused is None:
                exchange = "graph"
        self.exchange = exchange if (multi or self.fused is not None) else "none"
        self.borrow = bool(borrow_inputs)
        self.static = Batch(num_graphs=example.num_graphs, lut_col=example.lut_col,
                            sym_by_src=getattr(example, "sym_by_src", False),
                            **{k: ((getattr(example, k) if self.borrow else getattr(example, k).clone())
                                   if getattr(example, k) is not None else None) for k in _FIELDS})
        self.ptrs = {k: getattr(example, k).data_ptr() for k in _FIELDS if getattr(example, k) is not None}
        self.static.lut_rows = getattr(example, "lut_rows", None)
        self.shapes = {k: tuple(getattr(example, k).shape) for k in _FIELDS if getattr(example, k) is not None}
        # largest graph of the batch: sizes the shared memory of the block-per-graph kernels; read once here
        # (one host sync for a foreign batch), a constant of the captured step afterwards
        from . import ops
        self.static.max_nodes, self.static.max_edges = ops.batch_max_sizes(example)
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(torch.cuda.current_stream())
        # ---- snapshot: the warm-up below takes real optimizer steps
        model_sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        opt_sd = copy.deepcopy(optimizer.state_dict())
        if self.fused is not None:
            had_state = bool(self.fused.had_state or int(self.fused.state[1].item()) != 0)
        else:
            had_state = len(optimizer.state) > 0
        with torch.cuda.stream(self.stream):
            for _ in range(max(warmup, 1)):          # allocates workspaces / optimizer state eagerly
                self._front()
                self._tail()
            torch.cuda.synchronize()
            self._capture()
            torch.cuda.synchronize()
            # ---- restore in place (the graphs hold the addresses of these tensors)
            with torch.no_grad():
                for k, v in model.state_dict().items():
                    v.copy_(model_sd[k])
                if had_state:
                    cur = optimizer.state_dict()["state"]
                    for idx, st in opt_sd["state"].items():
                        for name, val in st.items():
                            if torch.is_tensor(val):
                                cur[idx][name].copy_(val)
                elif self.fused is not None:
                    self.fused.mark_fresh()          # next step: buffer = gradient, torch's rule for a fresh optimizer
                else:
                    # fresh optimizer: zeroed buffers reproduce the first-step rule of SGD / Adam exactly
                    # (buf = 0 * momentum + grad), except for SGD with dampening != 0
                    for group in optimizer.param_groups:
                        if group.get("dampening", 0) != 0:
                            raise NotImplementedError("GraphedTrainStep: SGD dampening != 0 with a fresh optimizer")
                    for st in optimizer.state.values():
                        for name, val in st.items():
                            if torch.is_tensor(val):
                                val.zero_()
            torch.cuda.synchronize()
        torch.cuda.current_stream().wait_stream(self.stream)

    # ------------------------------------------------------------------ capture
    def _capture(self) -> None:
        self.hyper = _hyper(self.opt)
        self.graph_b = None
        if self.exchange == "eager":
            self.graph_a = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_a, stream=self.stream, pool=self.pool, capture_error_mode="thread_local"):
                self.loss = self._front()
            self._capture_update()
        else:
            # one graph: zero -> forward -> loss -> backward (-> gather) -> exchange + optimizer
            self.graph_a = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_a, stream=self.stream, pool=self.pool, capture_error_mode="thread_local"):
                self.loss = self._front()
                self._tail()

    def _capture_update(self) -> None:
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_b, stream=self.stream, capture_error_mode="thread_local"):
            self.opt.step()

    def _front(self) -> torch.Tensor:
        self.static._cache = {}                      # the CSR of the batch is rebuilt inside the step
        if self.fused is not None:
            self._grads.zero()
        else:
            (self.ddp or self.opt).zero_grad(set_to_none=True)   # backward writes fresh gradients
        if self.loss_of is not None:
            loss = self.loss_of(self.ddp or self.model, self.static)
        else:
            loss = self.crit((self.ddp or self.model)(self.static), self.target_of(self.static))
        loss.backward()
        if self.fused is not None:
            self._grads.gather()                     # one concatenation kernel; p.grad -> flat views
        elif self.ddp is not None:
            self.ddp.grads.gather()
        return loss.detach()

    def _tail(self) -> None:
        """Gradient exchange + optimizer update."""
        if self.fused is not None:
            self.fused.step()
        else:
            self._exchange()
            self.opt.step()

    def _exchange(self) -> None:
        if self.ddp is not None:
            self.ddp.grads.all_reduce_mean()         # flat NCCL all-reduce (captured or eager)

    def _check_hyper(self) -> None:
        """A scheduler (``StepLR.step()``) changes ``param_groups``; the captured update holds the old
        values as host scalars: re-capture it (the whole step when it is one graph)."""
        if self.fused is not None:
            with torch.cuda.stream(self.stream):
                self.fused.sync_hyper()              # a few bytes to the device; the captured kernel reads them there
            return
        if _hyper(self.opt) == self.hyper:
            return
        with torch.cuda.stream(self.stream):
            torch.cuda.synchronize()
            self.hyper = _hyper(self.opt)
            if self.graph_b is not None:
                self._capture_update()
            else:
                # re-capturing runs nothing: parameters and optimizer state are untouched
                self.graph_a = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_a, stream=self.stream, capture_error_mode="thread_local"):
                    self.loss_new = self._front()
                    self._tail()
                self.loss = self.loss_new

    def step(self, batch: Batch) -> torch.Tensor:
        """Copies ``batch`` into the static buffers and replays the captured step; returns the
        (device) loss of this step -- reading it on the host is the caller's only sync."""
        for k, shape in self.shapes.items():
            t = getattr(batch, k)
            if t is None or tuple(t.shape) != shape:
                raise RuntimeError(f"GraphedTrainStep was captured for {k}{shape}; got "
                                   f"{None if t is None else tuple(t.shape)} (static shapes only)")
        mn, me = getattr(batch, "max_nodes", None), getattr(batch, "max_edges", None)
        if mn is None or me is None:
            from . import ops
            mn, me = ops.batch_max_sizes(batch)      # foreign batch: one device->host read, cached on the batch
        if mn > self.static.max_nodes or me > self.static.max_edges:
            raise RuntimeError(f"GraphedTrainStep was captured for graphs of <= {self.static.max_nodes} nodes / "
                               f"{self.static.max_edges} edges; got {mn} / {me}")
        self._check_hyper()
        if self.borrow and any(getattr(batch, k).data_ptr() != p for k, p in self.ptrs.items()):
            raise RuntimeError("GraphedTrainStep(borrow_inputs=True) replays on the tensors it was captured on; "
                               "this batch lives elsewhere")
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            if not self.borrow:
                for k in self.shapes:
                    getattr(self.static, k).copy_(getattr(batch, k), non_blocking=True)
            self.graph_a.replay()
            if self.graph_b is not None:
                self._exchange()
                self.graph_b.replay()
        torch.cuda.current_stream().wait_stream(self.stream)
        return self.loss


class GraphedStepCache:
    """CUDA-graphed training steps for RAGGED batches: one captured graph per distinct batch shape.

    ``lightpath_training/train.py:77-132`` walks its training set in fixed chunks with ``shuffle=False``, so the same
    batches -- the same (N, E, B, L) -- come back every ``num_chunks`` epochs; the step itself is ~45 small launches
    (0.3 ms of GPU work inside ~1 ms of Python and launch overhead per 512-graph batch).  The first time a shape is
    seen its step is captured (a few eager warm-up steps whose effect on the model and the optimizer is rolled back,
    then the capture); every later batch of that shape is one graph replay.  All captures share one memory pool
    (only one replays at a time), one flat gradient buffer and one fused optimizer tail, so the optimizer state is
    the same whichever graph runs.  Needs batches that carry their row counts on the host (``lut_rows``, set by
    ``PackedGraphStore.collate``) -- a device->host read inside a capture is an error."""

    def __init__(self, model, optimizer, loss_of: Callable, ddp=None, warmup: int = 2, max_entries: int = 4096,
                 borrow_inputs: bool = False):
        from .distributed import FlatGradBuffer, FusedSGDStep
        self.model, self.opt, self.loss_of, self.ddp = model, optimizer, loss_of, ddp
        self.warmup, self.max_entries = warmup, max_entries
        # borrow_inputs: key by the batch's STORAGE as well and capture on its own tensors -- a resident dataset whose
        # batch objects come back unchanged is then replayed without a single input copy
        self.borrow = bool(borrow_inputs)
        self.entries = {}
        self.pool = torch.cuda.graph_pool_handle()
        self.shared = None
        if FusedSGDStep.supports(optimizer):
            grads = ddp.grads if ddp is not None else FlatGradBuffer(model.parameters())
            self.shared = (grads, FusedSGDStep(optimizer, grads, distributed=ddp is not None))
        self.captures = self.replays = 0

    @staticmethod
    def key_of(batch) -> tuple:
        return tuple((k, tuple(getattr(batch, k).shape)) for k in _FIELDS if getattr(batch, k) is not None) + \
            (getattr(batch, "lut_rows", None), getattr(batch, "max_nodes", None), getattr(batch, "max_edges", None))

    def step(self, batch: Batch) -> torch.Tensor:
        key = self.key_of(batch)
        if self.borrow:
            key = key + tuple(getattr(batch, k).data_ptr() for k in _FIELDS if getattr(batch, k) is not None)
        g = self.entries.get(key)
        if g is None:
            if len(self.entries) >= self.max_entries:
                self.entries.pop(next(iter(self.entries)))
            g = GraphedTrainStep(self.model, self.opt, None, batch, ddp=self.ddp, warmup=self.warmup,
                                 loss_of=self.loss_of, shared=self.shared, pool=self.pool,
                                 exchange=None if self.shared is not None else "graph", borrow_inputs=self.borrow)
            self.entries[key] = g
            self.captures += 1
        self.replays += 1
        return g.step(batch)
