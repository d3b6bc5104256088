"""``Data`` / ``Batch`` containers with the PyG behaviour the reference relies on."""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

import torch


class Data:
    """Attribute bag for one graph.  Assigning ``None`` removes the key (PyG semantics that
    ``topological_training/dataset.py:107`` uses: ``data.x = None`` -> ``data.x`` reads ``None`` and
    the collate skips it)."""

    def __init__(self, **kwargs: Any):
        object.__setattr__(self, "_store", {})
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __setattr__(self, key: str, value: Any) -> None:
        if key.startswith("_"):
            object.__setattr__(self, key, value)
        elif value is None:
            self._store.pop(key, None)
        else:
            self._store[key] = value

    def __getattr__(self, key: str) -> Any:
        store = object.__getattribute__(self, "_store")
        if key in store:
            return store[key]
        if key in ("x", "edge_index", "edge_attr", "y", "batch", "ptr", "node_ids"):
            return None                                        # optional standard attributes
        raise AttributeError(key)

    def __contains__(self, key: str) -> bool:
        return key in self._store

    def keys(self) -> List[str]:
        return list(self._store)

    @property
    def num_nodes(self) -> int:
        s = self._store
        if "num_nodes" in s:
            return int(s["num_nodes"])
        if "x" in s:
            return int(s["x"].shape[0])
        if "batch" in s:
            return int(s["batch"].shape[0])
        if "node_ids" in s:
            return int(s["node_ids"].shape[0])
        ei = s.get("edge_index")
        return int(ei.max()) + 1 if ei is not None and ei.numel() else 0

    @num_nodes.setter
    def num_nodes(self, n: int) -> None:
        self._store["num_nodes"] = int(n)

    @property
    def num_edges(self) -> int:
        ei = self._store.get("edge_index")
        return int(ei.shape[1]) if ei is not None else 0

    def to(self, device, non_blocking: bool = False):
        out = self.__class__()
        for k, v in self._store.items():
            out._store[k] = v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v
        return out

    def cpu(self):
        return self.to("cpu")

    def __repr__(self) -> str:
        parts = [f"{k}={list(v.shape) if torch.is_tensor(v) else v!r}" for k, v in self._store.items()]
        return f"{self.__class__.__name__}({', '.join(parts)})"


class Batch(Data):
    """``Batch.from_data_list``: keys containing 'index' are concatenated on the last
    dimension and offset by the running node count, every other tensor on dimension 0; adds ``batch``,
    ``ptr``, ``num_graphs`` -- and ``edge_ptr`` (edge offsets per graph), which any in-order collate
    knows for free and the fused inference kernel uses."""

    @classmethod
    def from_data_list(cls, data_list: Sequence[Data]) -> "Batch":
        out = cls()
        B = len(data_list)
        n = [d.num_nodes for d in data_list]
        ptr = torch.zeros(B + 1, dtype=torch.int64)
        if B:
            ptr[1:] = torch.cumsum(torch.tensor(n, dtype=torch.int64), 0)
        keys: List[str] = []
        for d in data_list:
            for k in d.keys():
                if k not in keys:
                    keys.append(k)
        for k in keys:
            vals = [d._store.get(k) for d in data_list]
            if k == "num_nodes":
                continue
            if any(v is None for v in vals):
                continue                                       # PyG requires the key in every graph
            if all(torch.is_tensor(v) for v in vals):
                if "index" in k:
                    out._store[k] = torch.cat([v + int(ptr[i]) for i, v in enumerate(vals)], dim=-1)
                else:
                    vals = [v.unsqueeze(0) if v.dim() == 0 else v for v in vals]
                    out._store[k] = torch.cat(vals, dim=0)
            else:
                out._store[k] = list(vals)
        out._store["batch"] = torch.repeat_interleave(torch.arange(B, dtype=torch.int64), torch.tensor(n, dtype=torch.int64)) \
            if B else torch.zeros(0, dtype=torch.int64)
        out._store["ptr"] = ptr
        eptr = torch.zeros(B + 1, dtype=torch.int64)
        if B:
            eptr[1:] = torch.cumsum(torch.tensor([d.num_edges for d in data_list], dtype=torch.int64), 0)
        out._store["edge_ptr"] = eptr
        out._store["num_nodes"] = int(ptr[-1])
        object.__setattr__(out, "_num_graphs", B)
        return out

    @property
    def num_graphs(self) -> int:
        return int(object.__getattribute__(self, "_num_graphs")) if hasattr(self, "_num_graphs") else \
            (int(self._store["batch"].max()) + 1 if self._store.get("batch") is not None and self._store["batch"].numel() else 0)

    def to(self, device, non_blocking: bool = False):
        out = super().to(device, non_blocking)
        object.__setattr__(out, "_num_graphs", self.num_graphs)
        return out
