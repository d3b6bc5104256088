"""Host-facing streaming inference: pinned HOST batches in, host results out.

This is the end-to-end form of ``lightpath_training/test.py:77-94`` (``data.to(device)`` ->
``model(data)`` -> ``.cpu()`` per batch) without the per-batch stalls: ``depth`` slots, each with
its own CUDA stream, device staging buffers and pinned result buffers, so the H2D copy of batch
k+1 overlaps the kernel of batch k and the D2H of batch k-1.  One native call per batch
(``qot_lightpath_infer_host``) enqueues the copies, the fused eval kernel and the read-back; the
source row of ``edge_index`` is never copied -- the kernel reads the handful of entries it needs
straight from the pinned host buffer.  Nothing here computes on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Tuple

import torch

from . import _lib, ops


def _host_ptr(t: torch.Tensor, dtype, what: str) -> int:
    if t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"LightpathInferencePipeline: {what} must be a contiguous host {dtype} tensor")
    return t.data_ptr()


def _check_pinned(hb) -> None:
    """Once per batch object: every tensor the native call reads must be pinned host memory."""
    if getattr(hb, "_pinned_ok", False):
        return
    for k in ("x", "edge_index", "ptr", "edge_ptr", "lut_ptr"):
        t = getattr(hb, k)
        base = t._base if t._base is not None else t
        if t.is_cuda or not base.is_pinned():
            raise RuntimeError(f"LightpathInferencePipeline: {k} must be in pinned host memory "
                               "(Batch.pin_memory() / PackedGraphStore.host_batch(pin=True))")
    hb._pinned_ok = True


class _Slot:
    def __init__(self, dev, max_nodes, max_edges, max_graphs):
        self.stream = torch.cuda.Stream(device=dev)
        self.x = torch.empty(max_nodes, 5, dtype=torch.float32, device=dev)
        self.edst = torch.empty(max_edges, dtype=torch.int64, device=dev)
        self.ptrs = torch.empty(3 * (max_graphs + 1), dtype=torch.int64, device=dev)
        self.res = ops.new_infer_out(max_nodes, dev)
        self.z = torch.empty(_lib.lib().qot_lightpath_infer_workspace_bytes(max_nodes), dtype=torch.uint8, device=dev)
        self.arena = torch.empty(8 * max_edges + 24 * (max_graphs + 1) + 20 * max_nodes + 16, dtype=torch.uint8, device=dev)
        self.c = _lib.QotLpSlot(self.x.data_ptr(), None, self.edst.data_ptr(), self.ptrs.data_ptr(),
                                self.res.out.data_ptr(), self.res.lut_batch.data_ptr(),
                                self.res.lut_node.data_ptr(), self.res.n_lut.data_ptr(),
                                self.res.status.data_ptr(), self.z.data_ptr(), self.arena.data_ptr(), max_nodes, max_edges, max_graphs)
        self.out_h = torch.empty(max_nodes, 3, dtype=torch.float32).pin_memory()
        self.lb_h = torch.empty(max_nodes, dtype=torch.int64).pin_memory()
        self.st_h = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.done = torch.cuda.Event()
        self.rows = 0
        self.busy = False
        self.batch = None                         # keeps the host batch alive while the kernel reads it


class LightpathInferencePipeline:
    wire_note = ("copied in ONE transfer per step (the host batch keeps them contiguous): destination row of "
                 "edge_index, ptr/edge_ptr/lut_ptr, x; the source row stays in pinned host memory and the kernel "
                 "reads ~4 sectors (32 B) per LUT row from it over PCIe (estimated, included in h2d_bytes_per_step)")

    def __init__(self, model, max_nodes: int, max_edges: int, max_graphs: int, depth: int = 3):
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("LightpathInferencePipeline needs the model on a CUDA device (no CPU path)")
        if model.training:
            raise RuntimeError("LightpathInferencePipeline runs the eval-mode forward: call model.eval()")
        self.model, self.dev = model, p.device
        self.caps = (int(max_nodes), int(max_edges), int(max_graphs))
        self.slots = [_Slot(self.dev, *self.caps) for _ in range(depth)]
        self.h2d_bytes = self.d2h_bytes = self.zero_copy_bytes = self.steps = 0
        self._h2d, self._d2h = C.c_int64(0), C.c_int64(0)
        self.prepared = model.prepared()          # folded parameters, before any slot stream uses them
        torch.cuda.synchronize(self.dev)

    # -- one batch in flight -------------------------------------------------------
    def _submit(self, slot: _Slot, hb) -> None:
        N, E, B = hb.num_nodes, hb.num_edges, hb.num_graphs
        if hb.ptr is None or hb.edge_ptr is None or hb.lut_ptr is None or hb.lut_col != self.model.is_lut_index:
            raise RuntimeError("LightpathInferencePipeline needs batches carrying ptr, edge_ptr and lut_ptr "
                               "(PackedGraphStore.host_batch / collate provide them)")
        _check_pinned(hb)
        L = int(hb.lut_ptr[-1])                     # known on the host: no D2H of the row count
        if L == 0:
            raise ValueError("No LUT node found in the batch.")
        lib = _lib.lib()
        _lib.check(lib.qot_lightpath_infer_host(
            _host_ptr(hb.x, torch.float32, "x"), _host_ptr(hb.edge_index, torch.int64, "edge_index"), E,
            _host_ptr(hb.ptr, torch.int64, "ptr"), _host_ptr(hb.edge_ptr, torch.int64, "edge_ptr"),
            _host_ptr(hb.lut_ptr, torch.int64, "lut_ptr"), N, B, self.prepared.data_ptr(),
            int(self.model.is_lut_index), C.byref(slot.c), slot.out_h.data_ptr(), slot.lb_h.data_ptr(),
            slot.st_h.data_ptr(), C.byref(self._h2d), C.byref(self._d2h), slot.stream.cuda_stream),
            "qot_lightpath_infer_host")
        slot.done.record(slot.stream)
        self.h2d_bytes += self._h2d.value
        self.d2h_bytes += self._d2h.value
        self.zero_copy_bytes += 32 * 4 * L          # ~4 in-edges per LUT row, one 32 B sector each (estimate)
        slot.rows, slot.busy, slot.batch = L, True, hb
        self.steps += 1

    def _collect(self, slot: _Slot) -> Tuple[torch.Tensor, torch.Tensor]:
        slot.done.synchronize()
        slot.busy, slot.batch = False, None
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
