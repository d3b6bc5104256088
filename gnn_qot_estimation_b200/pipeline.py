"""Host-facing streaming inference: pinned HOST batches in, host results out.

This is the end-to-end form of ``lightpath_training/test.py:77-94`` (``data.to(device)`` ->
``model(data)`` -> ``.cpu()`` per batch) without the per-batch stalls: ``depth`` slots, each with
its own CUDA stream, device staging buffers and pinned result buffers, so the H2D copy of batch
k+1 overlaps the kernels of batch k and the D2H of batch k-1.  One native call per batch
(``qot_lightpath_infer_wire_host``) enqueues ONE host->device copy of the batch in the compact wire
format (:class:`~.batch.WireBatch`: int32 offsets, the four non-flag fp32 features, uint8 LUT
positions and graph-local destinations, no source row -- 0.65 KB per 32-node graph), the device-side unpack into the reference layout, the
persistent eval kernel (the same kernel resident batches take: bit-identical rows) and the
read-back.  Nothing here computes on the host.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Iterable, List, Tuple

import torch

from . import _lib
from .batch import WireBatch


class _Slot:
    def __init__(self, dev, max_nodes, max_edges, max_graphs):
        L = _lib.lib()
        self.stream = torch.cuda.Stream(device=dev)
        self.arena = torch.empty(int(L.qot_lightpath_wire_bytes(max_nodes, max_edges, max_graphs, max_nodes)), dtype=torch.uint8, device=dev)
        self.x = torch.empty(max(max_nodes, 1), 5, dtype=torch.float32, device=dev)
        self.edge_index = torch.empty(2, max(max_edges, 1), dtype=torch.int64, device=dev)
        self.ptrs = torch.empty(3 * (max_graphs + 1), dtype=torch.int64, device=dev)
        rows = max(max_nodes, 1)
        self.result = torch.zeros(int(L.qot_lightpath_wire_result_bytes(rows)), dtype=torch.uint8, device=dev)
        self.lut_node = torch.empty(rows, dtype=torch.int32, device=dev)
        self.n_lut = torch.zeros(1, dtype=torch.int32, device=dev)
        self.z = torch.empty(rows, 20, dtype=torch.float32, device=dev)
        d = _lib.QotLpBatch()
        d.lut_node, d.n_lut, d.z = self.lut_node.data_ptr(), self.n_lut.data_ptr(), self.z.data_ptr()
        self.desc = torch.frombuffer(bytearray(bytes(d)), dtype=torch.uint8).to(dev)
        self.c = _lib.QotLpWireSlot(self.arena.data_ptr(), self.x.data_ptr(), self.edge_index.data_ptr(), self.ptrs.data_ptr(),
                                    self.desc.data_ptr(), self.result.data_ptr(), self.lut_node.data_ptr(),
                                    self.n_lut.data_ptr(), max_nodes, max_edges, max_graphs)
        self.done = torch.cuda.Event()
        self.rows = 0
        self.busy = False
        self.batch = None                         # keeps the host arena alive while the copy engine reads it
        self.region = None                        # (offset, bytes) of this batch's results in the pipeline's host buffer


class LightpathInferencePipeline:
    wire_note = ("ONE host->device copy per step of the batch in the compact wire format (int32 ptr / edge_ptr / "
                 "lut_ptr, the 4 non-flag fp32 node features, uint8 LUT positions and graph-local destination ids; no source row: "
                 "verified from_networkx layout), "
                 "unpacked to the reference layout on the device")

    def __init__(self, model, max_nodes: int, max_edges: int, max_graphs: int, depth: int = 6):
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("LightpathInferencePipeline needs the model on a CUDA device (no CPU path)")
        if model.training:
            raise RuntimeError("LightpathInferencePipeline runs the eval-mode forward: call model.eval()")
        model._check_supported()
        self.model, self.dev = model, p.device
        self.caps = (int(max_nodes), int(max_edges), int(max_graphs))
        with torch.cuda.device(self.dev):
            self.slots = [_Slot(self.dev, *self.caps) for _ in range(depth)]
            self.h2d_bytes = self.d2h_bytes = self.steps = 0
            self.zero_copy_bytes = 0                  # (round-1 field: nothing is read over PCIe by the kernel any more)
            self._h2d, self._d2h = C.c_int64(0), C.c_int64(0)
            self._res = self._res_cache = None
            self.prepared = model.prepared()          # folded parameters, before any slot stream uses them
            torch.cuda.synchronize(self.dev)

    # -- one batch in flight -------------------------------------------------------
    def _submit(self, slot: _Slot, wb: WireBatch, off: int) -> None:
        lib = _lib.lib()
        _lib.check(lib.qot_lightpath_infer_wire_host(
            wb.arena.data_ptr(), wb.num_nodes, wb.num_edges, wb.num_graphs, wb.rows, self.prepared.data_ptr(),
            int(self.model.is_lut_index), C.byref(slot.c), self._res.data_ptr() + off,
            C.byref(self._h2d), C.byref(self._d2h), slot.stream.cuda_stream),
            "qot_lightpath_infer_wire_host")
        slot.done.record(slot.stream)
        self.h2d_bytes += self._h2d.value
        self.d2h_bytes += self._d2h.value
        slot.rows, slot.busy, slot.batch, slot.region = wb.rows, True, wb, off
        self.steps += 1

    def _collect(self, slot: _Slot) -> Tuple[torch.Tensor, torch.Tensor]:
        slot.done.synchronize()
        slot.busy, slot.batch = False, None
        off, n = slot.region, slot.rows
        status = int(self._res[off:off + 4].view(torch.int32)[0])
        if status != 0:
            raise RuntimeError("LightpathInferencePipeline: a batch's offsets / lut_ptr do not describe its contents "
                               f"(device status {status})")
        o0 = off + 16
        o1 = o0 + ((12 * n + 15) & ~15)
        return self._res[o0:o0 + 12 * n].view(torch.float32).view(n, 3), self._res[o1:o1 + 8 * n].view(torch.int64)

    def _check(self, wb) -> None:
        if not isinstance(wb, WireBatch):
            raise RuntimeError("LightpathInferencePipeline takes WireBatch objects (PackedGraphStore.host_wire_batch)")
        if wb.lut_col != self.model.is_lut_index:
            raise RuntimeError("LightpathInferencePipeline: the batch was packed for another is_lut_index")
        if wb.arena.is_cuda or not wb.arena.is_pinned():
            raise RuntimeError("LightpathInferencePipeline: the batch arena must be pinned host memory "
                               "(PackedGraphStore.host_wire_batch(pin=True))")
        if wb.rows == 0:
            raise ValueError("No LUT node found in the batch.")

    def run(self, host_batches: Iterable[WireBatch]) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """Returns ``[(out [L,3], lut_batch [L]), ...]`` as HOST tensors, one per batch, in order.  The results of a
        run live in ONE pinned buffer the device copies straight into (views, no per-batch clone); the buffer belongs
        to the returned tensors (a later run allocates its own when they are still alive)."""
        batches = list(host_batches)
        lib = _lib.lib()
        offs, tot = [], 0
        for wb in batches:
            self._check(wb)
            offs.append(tot)
            tot += (int(lib.qot_lightpath_wire_result_bytes(wb.rows)) + 15) & ~15
        buf = self._res_cache
        if buf is None or buf.numel() < tot or sys.getrefcount(buf) > 3:      # still referenced by earlier results
            buf = torch.empty(max(tot, 16), dtype=torch.uint8).pin_memory()
        self._res, self._res_cache = buf, buf
        results = []
        depth = len(self.slots)
        with torch.cuda.device(self.dev):
            for k, wb in enumerate(batches):
                slot = self.slots[k % depth]
                if slot.busy:
                    results.append(self._collect(slot))
                self._submit(slot, wb, offs[k])
            k = len(batches)
            for j in range(max(k - depth, 0), k):
                slot = self.slots[j % depth]
                if slot.busy:
                    results.append(self._collect(slot))
        self._res = None
        return results
