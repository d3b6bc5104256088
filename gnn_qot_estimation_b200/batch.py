"""Batch container + packed graph store feeding the device-side collate.

``Batch`` carries exactly the attributes the reference modules read from a PyG
batch (topological_training/models.py:44-52, lightpath_training/models.py:27,35;
``num_graphs`` at topological_training/train.py:117), plus the offsets PyG keeps
as ``ptr`` and the edge offsets our collate knows for free (``edge_ptr``).

``PackedGraphStore`` replaces one-pickle-per-graph + ``Dataset.__getitem__`` +
PyG's collate: all graphs live in flat
arrays in HBM and ``collate`` is one kernel launch (csrc/graph_index.cu).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _lib

_FIELDS = ("x", "edge_index", "edge_attr", "batch", "node_ids", "y", "ptr", "edge_ptr", "lut_ptr")


class Batch:
    """Minimal stand-in for ``torch_geometric.data.Batch``."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, batch=None, node_ids=None,
                 y=None, ptr=None, edge_ptr=None, lut_ptr=None, num_graphs: Optional[int] = None,
                 lut_col: Optional[int] = None, sym_by_src: bool = False):
        self.x, self.edge_index, self.edge_attr = x, edge_index, edge_attr
        self.batch, self.node_ids, self.y = batch, node_ids, y
        self.ptr, self.edge_ptr = ptr, edge_ptr
        # lut_ptr [B+1]: exclusive prefix of the per-graph count of nodes with x[:, lut_col] == 1.0
        # (where each graph's readout rows go); built by the collate like ptr / edge_ptr
        self.lut_ptr, self.lut_col = lut_ptr, lut_col
        self.num_graphs = num_graphs
        # True only when the producer of the batch VERIFIED the from_networkx layout of undirected graphs for every
        # graph (edges grouped by source ascending, both directions present, no duplicate pair):
        # PackedGraphStore.verify_layout().  Lets the eval kernels skip the source row of edge_index.
        self.sym_by_src = bool(sym_by_src)
        # number of readout rows (lut_ptr[-1]) when the producer of the batch knows it on the HOST (collates of a
        # PackedGraphStore do): the training path then needs no device->host read and can be captured in a CUDA graph
        self.lut_rows = None
        self._cache = {}          # per-batch CSR etc. built lazily by the ops layer

    # -- PyG-like conveniences ------------------------------------------------
    @property
    def num_nodes(self) -> int:
        if self.batch is not None:
            return int(self.batch.shape[0])
        if self.x is not None:
            return int(self.x.shape[0])
        return int(self.node_ids.shape[0])

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])

    def to(self, device, non_blocking: bool = False) -> "Batch":
        kw = {k: (getattr(self, k).to(device, non_blocking=non_blocking)
                  if getattr(self, k) is not None else None) for k in _FIELDS}
        b = Batch(num_graphs=self.num_graphs, lut_col=self.lut_col, sym_by_src=self.sym_by_src, **kw)
        b.lut_rows = self.lut_rows
        for k in ("max_nodes", "max_edges"):
            if hasattr(self, k):
                setattr(b, k, getattr(self, k))
        return b

    def pin_memory(self) -> "Batch":
        kw = {k: (getattr(self, k).pin_memory() if getattr(self, k) is not None else None)
              for k in _FIELDS}
        return Batch(num_graphs=self.num_graphs, lut_col=self.lut_col, sym_by_src=self.sym_by_src, **kw)

    def cpu(self) -> "Batch":
        return self.to("cpu")

    def nbytes(self, fields: Sequence[str] = _FIELDS) -> int:
        return sum(getattr(self, k).numel() * getattr(self, k).element_size()
                   for k in fields if getattr(self, k) is not None)

    def __repr__(self):
        parts = [f"{k}={tuple(getattr(self, k).shape)}" for k in _FIELDS if getattr(self, k) is not None]
        return f"Batch(num_graphs={self.num_graphs}, " + ", ".join(parts) + ")"


class WireBatch:
    """A batch in the compact wire format of include/qot_b200.h (qot_lightpath_infer_wire_host): ONE contiguous
    (pinned) host arena  int32 ptr | int32 edge_ptr | int32 lut_ptr | x4 fp32 [N,4] | uint8 lut_local [L] | uint8 dst [E]
    -- the node features without the LUT flag column (exactly 0.0 / 1.0: it travels as the graph-local position of
    every readout row), destinations as graph-local ids, no source row (verified from_networkx layout only).
    ~0.65 KB per 32-node graph instead of the 3.4 KB of the reference tensors (x + int64 edge_index [2,E] + offsets)."""

    def __init__(self, arena: torch.Tensor, N: int, E: int, B: int, L: int, lut_col: int):
        self.arena, self.num_nodes, self.num_edges, self.num_graphs, self.rows, self.lut_col = arena, N, E, B, L, lut_col

    @property
    def nbytes(self) -> int:
        return int(self.arena.numel())

    @staticmethod
    def offsets(N: int, E: int, B: int, L: int):
        a16 = lambda v: (v + 15) & ~15
        o_x = a16(12 * (B + 1))
        o_l = o_x + a16(16 * N)
        o_d = o_l + a16(L)
        return o_x, o_l, o_d, o_d + a16(E)


class PackedGraphStore:
    """All graphs of a dataset as flat arrays.

    node_ptr/edge_ptr [G+1] int64; edge_src/edge_dst [E_tot] int32 graph-local ids in
    ``from_networkx`` order (grouped by source); node_feat [N_tot,F]
    or None (topological: ``x=None`` -> embeddings, topological_training/dataset.py:107);
    edge_feat [E_tot,D] or None; y [G,3].
    """

    def __init__(self, node_ptr, edge_ptr, edge_src, edge_dst, node_feat=None, edge_feat=None, y=None,
                 lut_col: Optional[int] = None):
        self.lut_col = lut_col       # column of node_feat holding the LUT flag (lightpath graphs) or None
        self.node_ptr, self.edge_ptr = node_ptr.to(torch.int64), edge_ptr.to(torch.int64)
        self.edge_src, self.edge_dst = edge_src.to(torch.int32), edge_dst.to(torch.int32)
        self.node_feat, self.edge_feat, self.y = node_feat, edge_feat, y
        self._node_ptr_host = self.node_ptr.cpu().numpy()
        self._edge_ptr_host = self.edge_ptr.cpu().numpy()
        self._arange = None
        self.sym_by_src = False      # set by verify_layout()
        self._lut_count_host = None  # per-graph LUT-node counts (host), built once on first collate

    @property
    def num_graphs(self) -> int:
        return int(self._node_ptr_host.shape[0] - 1)

    @property
    def device(self):
        return self.node_ptr.device

    def to(self, device) -> "PackedGraphStore":
        mv = lambda t: None if t is None else t.to(device)
        st = PackedGraphStore(mv(self.node_ptr), mv(self.edge_ptr), mv(self.edge_src), mv(self.edge_dst),
                              mv(self.node_feat), mv(self.edge_feat), mv(self.y), self.lut_col)
        st.sym_by_src = self.sym_by_src
        return st

    def verify_layout(self) -> bool:
        """One-time check (on the store's device, a few sorts over the edge arrays -- not on the hot path) of the
        layout ``torch_geometric.utils.from_networkx`` gives an undirected ``nx.Graph`` (SURVEY.md A.6;
        lightpath_training/dataset.py:86): inside every graph the edges are grouped by source node ascending,
        every edge is present in both directions, no (source, destination) pair repeats, endpoints lie inside
        the graph.  Sets and returns ``sym_by_src``; batches collated from the store inherit it, and the
        LightpathGNN eval kernels then never read the source row of ``edge_index``."""
        E = int(self.edge_src.numel())
        ok = True
        if E:
            n = (self.node_ptr[1:] - self.node_ptr[:-1])
            cnt = (self.edge_ptr[1:] - self.edge_ptr[:-1])
            g = torch.repeat_interleave(torch.arange(self.num_graphs, device=self.device), cnt)
            src, dst = self.edge_src.to(torch.int64), self.edge_dst.to(torch.int64)
            ng = n[g]
            K = int(n.max().item()) + 1
            ok = bool(((src >= 0) & (src < ng) & (dst >= 0) & (dst < ng)).all().item())
            if ok:
                same = g[1:] == g[:-1]
                ok = bool((~same | (src[1:] >= src[:-1])).all().item())
            if ok:
                kf = (g * K + src) * K + dst
                kr = (g * K + dst) * K + src
                kf_sorted = torch.sort(kf).values
                ok = bool((kf_sorted[1:] != kf_sorted[:-1]).all().item()) and \
                    bool(torch.equal(kf_sorted, torch.sort(kr).values))
        self.sym_by_src = ok
        return ok

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.node_ptr, self.edge_ptr, self.edge_src, self.edge_dst, self.node_feat,
                    self.edge_feat, self.y) if t is not None)

    # -- device-side collate ---------------------------------------------------
    @_lib.on_tensor_device
    def collate(self, ids: Union[range, slice, torch.Tensor, Sequence[int]]) -> Batch:
        """Collates graphs ``ids`` into one :class:`Batch` on the store's CUDA device
        with one launch of ``qot_collate``; no device->host synchronisation."""
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("PackedGraphStore.collate runs on the GPU: move the store with .to('cuda')")
        lib = _lib.lib()
        nph, eph = self._node_ptr_host, self._edge_ptr_host
        if isinstance(ids, slice):
            ids = range(*ids.indices(self.num_graphs))
        if isinstance(ids, range) and ids.step == 1:
            g0, g1 = ids.start, ids.stop
            B = g1 - g0
            N, E = int(nph[g1] - nph[g0]), int(eph[g1] - eph[g0])
            if self._arange is None or self._arange.numel() < self.num_graphs:
                self._arange = torch.arange(self.num_graphs, dtype=torch.int64, device=dev)
            gids = self._arange[g0:g1]
            optr = self.node_ptr[g0:g1 + 1] - self.node_ptr[g0]
            oeptr = self.edge_ptr[g0:g1 + 1] - self.edge_ptr[g0]
        else:
            ids_np = np.asarray(ids.cpu() if isinstance(ids, torch.Tensor) else list(ids), dtype=np.int64)
            B = int(ids_np.shape[0])
            cn = np.zeros(B + 1, dtype=np.int64)
            ce = np.zeros(B + 1, dtype=np.int64)
            np.cumsum(nph[ids_np + 1] - nph[ids_np], out=cn[1:])
            np.cumsum(eph[ids_np + 1] - eph[ids_np], out=ce[1:])
            N, E = int(cn[-1]), int(ce[-1])
            packed = torch.from_numpy(np.concatenate([ids_np, cn, ce])).to(dev, non_blocking=True)
            gids, optr, oeptr = packed[:B], packed[B:2 * B + 1], packed[2 * B + 1:]
        F = self.node_feat.shape[1] if self.node_feat is not None else 0
        D = self.edge_feat.shape[1] if self.edge_feat is not None else 0
        Y = self.y.shape[1] if self.y is not None else 0
        x = torch.empty(N, F, dtype=torch.float32, device=dev) if F else None
        ei = torch.empty(2, E, dtype=torch.int64, device=dev)
        ea = torch.empty(E, D, dtype=torch.float32, device=dev) if D else None
        bt = torch.empty(N, dtype=torch.int64, device=dev)
        nid = torch.empty(N, dtype=torch.int64, device=dev) if not F else None
        y = torch.empty(B, Y, dtype=torch.float32, device=dev) if Y else None
        st = _lib.QotStore(_lib.ptr(self.node_ptr), _lib.ptr(self.edge_ptr), _lib.ptr(self.edge_src),
                           _lib.ptr(self.edge_dst), _lib.ptr(self.node_feat), _lib.ptr(self.edge_feat),
                           _lib.ptr(self.y), F, D, Y)
        optr = optr.contiguous()
        oeptr = oeptr.contiguous()
        _lib.check(lib.qot_collate(C.byref(st), _lib.ptr(gids.contiguous()), B, _lib.ptr(optr), _lib.ptr(oeptr),
                                   N, E, _lib.ptr(x), _lib.ptr(ei), _lib.ptr(ea), _lib.ptr(bt),
                                   _lib.ptr(nid), _lib.ptr(y), _lib.stream()), "qot_collate")
        lut_ptr = None
        if self.lut_col is not None and x is not None:
            from . import ops
            lut_ptr = ops.lightpath_lut_ptr(x, optr, self.lut_col)
        b = Batch(x=x, edge_index=ei, edge_attr=ea, batch=bt, node_ids=nid, y=y,
                  ptr=optr, edge_ptr=oeptr, lut_ptr=lut_ptr, num_graphs=B, lut_col=self.lut_col,
                  sym_by_src=self.sym_by_src)
        # largest graph of the batch (host arrays, no sync): sizes the per-graph kernels' shared memory
        if isinstance(ids, range) and ids.step == 1:
            dn, de = np.diff(nph[g0:g1 + 1]), np.diff(eph[g0:g1 + 1])
        else:
            dn, de = nph[ids_np + 1] - nph[ids_np], eph[ids_np + 1] - eph[ids_np]
        b.max_nodes = int(dn.max()) if B else 0
        b.max_edges = int(de.max()) if B else 0
        if lut_ptr is not None:
            if self._lut_count_host is None:             # once per store: one device->host read
                flags = (self.node_feat[:, self.lut_col] == 1.0).to(torch.int64)
                gid = torch.repeat_interleave(torch.arange(self.num_graphs, device=dev), self.node_ptr[1:] - self.node_ptr[:-1])
                self._lut_count_host = torch.zeros(self.num_graphs, dtype=torch.int64, device=dev).index_add_(0, gid, flags).cpu().numpy()
            c = self._lut_count_host
            b.lut_rows = int(c[g0:g1].sum()) if isinstance(ids, range) and ids.step == 1 else int(c[ids_np].sum())
        return b

    # -- the same range in the compact wire format (what travels to the GPU in LightpathInferencePipeline)
    def host_wire_batch(self, g0: int, g1: int, pin: bool = True) -> WireBatch:
        """Graphs [g0, g1) packed for qot_lightpath_infer_wire_host.  Needs a host-resident store whose layout has
        been verified (``verify_layout()``: the format carries no source row) and graphs of <= 255 nodes."""
        if self.device.type != "cpu":
            raise RuntimeError("host_wire_batch needs a host-resident store")
        if not self.sym_by_src:
            raise RuntimeError("host_wire_batch: call verify_layout() first -- the wire format has no source row and is "
                               "only defined for the verified from_networkx layout")
        if self.lut_col is None or self.node_feat is None or self.node_feat.shape[1] != 5:
            raise RuntimeError("host_wire_batch packs lightpath graphs (x [N,5] with a LUT flag column)")
        n0, n1 = int(self._node_ptr_host[g0]), int(self._node_ptr_host[g1])
        e0, e1 = int(self._edge_ptr_host[g0]), int(self._edge_ptr_host[g1])
        B, N, E = g1 - g0, n1 - n0, e1 - e0
        ptr = (self.node_ptr[g0:g1 + 1] - n0)
        if int((ptr[1:] - ptr[:-1]).max()) > 255:
            raise RuntimeError("host_wire_batch: a graph has more than 255 nodes (uint8 destination ids)")
        x = self.node_feat[n0:n1]
        flag = x[:, self.lut_col]
        is_lut = flag == 1.0
        if not bool((is_lut | (flag == 0.0)).all()):
            raise RuntimeError("host_wire_batch: the LUT flag column must hold exactly 0.0 / 1.0 (it is shipped as positions)")
        bt = torch.repeat_interleave(torch.arange(B, dtype=torch.int64), ptr[1:] - ptr[:-1])
        lut = torch.zeros(B + 1, dtype=torch.int64)
        torch.cumsum(torch.zeros(B, dtype=torch.int64).index_add_(0, bt, is_lut.to(torch.int64)), 0, out=lut[1:])
        L = int(lut[-1])
        rows = torch.nonzero(is_lut).view(-1)                         # readout rows in ascending node order
        o_x, o_l, o_d, nbytes = WireBatch.offsets(N, E, B, L)
        arena = torch.zeros(nbytes, dtype=torch.uint8)
        if pin:
            arena = arena.pin_memory()
        ptrs = arena[:12 * (B + 1)].view(torch.int32).view(3, B + 1)
        ptrs[0].copy_(ptr); ptrs[1].copy_(self.edge_ptr[g0:g1 + 1] - e0); ptrs[2].copy_(lut)
        cols = [c for c in range(5) if c != self.lut_col]
        arena[o_x:o_x + 16 * N].view(torch.float32).view(N, 4).copy_(x[:, cols])
        arena[o_l:o_l + L].copy_((rows - ptr[bt[rows]]).to(torch.uint8))
        arena[o_d:o_d + E].copy_(self.edge_dst[e0:e1].to(torch.uint8))
        return WireBatch(arena, N, E, B, L, self.lut_col)

    # -- host-side view of a contiguous range (what a host DataLoader would hand over)
    def host_batch(self, g0: int, g1: int, pin: bool = False) -> Batch:
        """Batch of graphs [g0,g1) as HOST tensors in the reference's layout (the
        buffers a user of the reference passes to ``data.to(device)``)."""
        if self.device.type != "cpu":
            raise RuntimeError("host_batch needs a host-resident store")
        n0, n1 = int(self._node_ptr_host[g0]), int(self._node_ptr_host[g1])
        e0, e1 = int(self._edge_ptr_host[g0]), int(self._edge_ptr_host[g1])
        B = g1 - g0
        ptr = self.node_ptr[g0:g1 + 1] - n0
        eptr = self.edge_ptr[g0:g1 + 1] - e0
        counts_e = (eptr[1:] - eptr[:-1])
        off = torch.repeat_interleave(ptr[:-1], counts_e)
        ei = torch.stack([self.edge_src[e0:e1].to(torch.int64) + off,
                          self.edge_dst[e0:e1].to(torch.int64) + off])
        counts_n = ptr[1:] - ptr[:-1]
        bt = torch.repeat_interleave(torch.arange(B, dtype=torch.int64), counts_n)
        x = self.node_feat[n0:n1].clone() if self.node_feat is not None else None
        nid = None
        if x is None:
            nid = torch.arange(n1 - n0, dtype=torch.int64) - torch.repeat_interleave(ptr[:-1], counts_n)
        ea = self.edge_feat[e0:e1].clone() if self.edge_feat is not None else None
        y = self.y[g0:g1].clone() if self.y is not None else None
        lut_ptr = None
        if self.lut_col is not None and x is not None:
            lut_ptr = torch.zeros(B + 1, dtype=torch.int64)
            flags = (x[:, self.lut_col] == 1.0).to(torch.int64)
            torch.cumsum(torch.zeros(B, dtype=torch.int64).index_add_(0, bt, flags), 0, out=lut_ptr[1:])
        if pin and lut_ptr is not None and x is not None:
            # one pinned arena per batch: [edge_index (src row | dst row) | ptr | edge_ptr | lut_ptr | x].  Everything
            # the fused eval kernel needs on the device -- the destination row, the three offset arrays and x --
            # is then ONE contiguous range: qot_lightpath_infer_host moves it with a single copy
            E_, N_ = int(ei.shape[1]), int(x.shape[0])
            nbytes = 16 * E_ + 24 * (B + 1) + 20 * N_
            arena = torch.empty(max(nbytes, 8), dtype=torch.uint8).pin_memory()
            ei_v = arena[:16 * E_].view(torch.int64).view(2, E_)
            ptrs = arena[16 * E_:16 * E_ + 24 * (B + 1)].view(torch.int64).view(3, B + 1)
            x_v = arena[16 * E_ + 24 * (B + 1):nbytes].view(torch.float32).view(N_, x.shape[1])
            ei_v.copy_(ei); ptrs[0].copy_(ptr); ptrs[1].copy_(eptr); ptrs[2].copy_(lut_ptr); x_v.copy_(x)
            pin_ = lambda t: None if t is None else t.pin_memory()
            b = Batch(x=x_v, edge_index=ei_v, edge_attr=pin_(ea), batch=pin_(bt), node_ids=pin_(nid), y=pin_(y),
                      ptr=ptrs[0], edge_ptr=ptrs[1], lut_ptr=ptrs[2], num_graphs=B, lut_col=self.lut_col,
                      sym_by_src=self.sym_by_src)
            b._arena = arena                              # keeps the pinned allocation alive
            return b
        # the three offset arrays share one [3, B+1] buffer: one H2D copy moves them all
        if lut_ptr is not None:
            ptrs = torch.stack([ptr, eptr, lut_ptr])
            if pin:
                ptrs = ptrs.pin_memory()
            ptr_v, eptr_v, lut_v = ptrs[0], ptrs[1], ptrs[2]
        else:
            ptr_v, eptr_v, lut_v = ptr.clone(), eptr.clone(), None
            if pin:
                ptr_v, eptr_v = ptr_v.pin_memory(), eptr_v.pin_memory()
        pin_ = (lambda t: None if t is None else t.pin_memory()) if pin else (lambda t: t)
        return Batch(x=pin_(x), edge_index=pin_(ei), edge_attr=pin_(ea), batch=pin_(bt), node_ids=pin_(nid),
                     y=pin_(y), ptr=ptr_v, edge_ptr=eptr_v, lut_ptr=lut_v, num_graphs=B, lut_col=self.lut_col,
                     sym_by_src=self.sym_by_src)
