"""Host-facing streaming inference: pinned HOST batches in, host results out.

This is the end-to-end form of ``lightpath_training/test.py:77-94`` (``data.to(device)`` ->
``model(data)`` -> ``.cpu()`` per batch) without the per-batch stalls: ``depth`` slots, each with
its own CUDA stream, device staging buffers and pinned result buffers, so the H2D copy of batch
k+1 overlaps the kernels of batch k and the D2H of batch k-1.  Nothing here computes: the
arithmetic is the fused eval kernel behind ``qot_lightpath_infer``.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch

from . import ops


class _Slot:
    def __init__(self, dev, max_nodes, max_edges, max_graphs):
        self.stream = torch.cuda.Stream(device=dev)
        self.x = torch.empty(max_nodes, 5, dtype=torch.float32, device=dev)
        self.ei = torch.empty(2 * max_edges, dtype=torch.int64, device=dev)
        self.gptr = torch.empty(max_graphs + 1, dtype=torch.int64, device=dev)
        self.eptr = torch.empty(max_graphs + 1, dtype=torch.int64, device=dev)
        self.lptr = torch.empty(max_graphs + 1, dtype=torch.int64, device=dev)
        self.res = ops.new_infer_out(max_nodes, dev)
        self.out_h = torch.empty(max_nodes, 3, dtype=torch.float32).pin_memory()
        self.lb_h = torch.empty(max_nodes, dtype=torch.int64).pin_memory()
        self.st_h = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.done = torch.cuda.Event()
        self.rows = 0
        self.busy = False


class LightpathInferencePipeline:
    def __init__(self, model, max_nodes: int, max_edges: int, max_graphs: int, depth: int = 3):
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("LightpathInferencePipeline needs the model on a CUDA device (no CPU path)")
        if model.training:
            raise RuntimeError("LightpathInferencePipeline runs the eval-mode forward: call model.eval()")
        self.model, self.dev = model, p.device
        self.caps = (int(max_nodes), int(max_edges), int(max_graphs))
        self.slots = [_Slot(self.dev, *self.caps) for _ in range(depth)]
        self.h2d_bytes = self.d2h_bytes = self.steps = 0
        model.prepared()                          # fold the parameters once, before any slot stream uses them
        torch.cuda.synchronize(self.dev)

    # -- one batch in flight -------------------------------------------------------
    def _submit(self, slot: _Slot, hb) -> None:
        N, E, B = hb.num_nodes, hb.num_edges, hb.num_graphs
        if N > self.caps[0] or E > self.caps[1] or B > self.caps[2]:
            raise RuntimeError(f"batch (N={N}, E={E}, B={B}) exceeds the pipeline capacity {self.caps}")
        if hb.ptr is None or hb.edge_ptr is None or hb.lut_ptr is None or hb.lut_col != self.model.is_lut_index:
            raise RuntimeError("LightpathInferencePipeline needs batches carrying ptr, edge_ptr and lut_ptr "
                               "(PackedGraphStore.host_batch / collate provide them)")
        L = int(hb.lut_ptr[-1])                     # known on the host: no D2H of the row count
        if L == 0:
            raise ValueError("No LUT node found in the batch.")
        with torch.cuda.stream(slot.stream):
            x = slot.x[:N]
            ei = slot.ei[: 2 * E].view(2, E)
            gptr, eptr, lptr = slot.gptr[: B + 1], slot.eptr[: B + 1], slot.lptr[: B + 1]
            x.copy_(hb.x, non_blocking=True)
            ei.copy_(hb.edge_index, non_blocking=True)
            gptr.copy_(hb.ptr, non_blocking=True)
            eptr.copy_(hb.edge_ptr, non_blocking=True)
            lptr.copy_(hb.lut_ptr, non_blocking=True)
            self.h2d_bytes += 20 * N + 16 * E + 24 * (B + 1)
            ops.lightpath_infer(x, ei, gptr, eptr, lptr, self.model.prepared(), self.model.is_lut_index, slot.res)
            slot.st_h.copy_(slot.res.status, non_blocking=True)
            slot.out_h[:L].copy_(slot.res.out[:L], non_blocking=True)
            slot.lb_h[:L].copy_(slot.res.lut_batch[:L], non_blocking=True)
            self.d2h_bytes += 4 + 20 * L
            slot.rows = L
            slot.done.record(slot.stream)
        slot.busy = True
        self.steps += 1

    def _collect(self, slot: _Slot) -> Tuple[torch.Tensor, torch.Tensor]:
        slot.done.synchronize()
        slot.busy = False
        if int(slot.st_h[0]) != 0:
            raise RuntimeError("LightpathInferencePipeline: a batch's lut_ptr does not match its x")
        n = slot.rows
        return slot.out_h[:n].clone(), slot.lb_h[:n].clone()

    def run(self, host_batches: Iterable) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """Returns ``[(out [L,3], lut_batch [L]), ...]`` as HOST tensors, one per batch, in order."""
        results = []
        depth = len(self.slots)
        k = 0
        for hb in host_batches:
            slot = self.slots[k % depth]
            if slot.busy:
                results.append(self._collect(slot))
            self._submit(slot, hb)
            k += 1
        for j in range(max(k - depth, 0), k):
            slot = self.slots[j % depth]
            if slot.busy:
                results.append(self._collect(slot))
        return results
