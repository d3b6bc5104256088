"""B200-native message-passing hot path of santiagolmedo/gnn_qot_estimation.

Drop-in modules (same constructors, ``forward(data)`` and state_dict names as the
reference): :class:`TopologicalGNN` (topological_training/models.py) and
:class:`LightpathGNN` (lightpath_training/models.py), running on hand-written
sm_100a kernels behind the C ABI in ``include/qot_b200.h``.
"""
from .batch import Batch, PackedGraphStore  # noqa: F401

__all__ = ["Batch", "PackedGraphStore", "TopologicalGNN", "LightpathGNN"]


def __getattr__(name):
    if name == "TopologicalGNN":
        from .topological_training.models import TopologicalGNN
        return TopologicalGNN
    if name == "LightpathGNN":
        from .lightpath_training.models import LightpathGNN
        return LightpathGNN
    raise AttributeError(name)
