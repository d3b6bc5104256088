"""``torch_geometric.loader.DataLoader``: a torch DataLoader whose collate is Batch.from_data_list
(call sites topological_training/train.py:72-74,93-95, lightpath_training/test.py:36)."""
from __future__ import annotations

import torch.utils.data

from .data import Batch


def _collate(items):
    return Batch.from_data_list(list(items))


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle, collate_fn=_collate, **kwargs)
