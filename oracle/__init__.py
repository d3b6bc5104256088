"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the QoT message-passing hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the reported CPU
baseline -- never as the thing shipped.  The product path
(``gnn_qot_estimation_b200``) never imports this package and fails loudly when
its CUDA library is missing.

PARITY UNPINNED: the arithmetic of the reference hot path lives in PyTorch
Geometric (``torch_geometric``, version not pinned by the reference; checkpoint
key layout implies >= 2.5), which is not installed here and is not vendored under
``/root/reference``.  The reference ships no tests, golden vectors or
known-answer fixtures for this path.  The oracle therefore restates PyG's
*published* layer semantics and is anchored on
  * strict ``load_state_dict`` of the three shipped checkpoints,
  * hand-derived micro known-answer cases (tests/test_oracle_kat.py),
  * fp64 gradcheck and direct-vs-factorised NNConv identities.
"""
from .qot_oracle import *  # noqa: F401,F403
from .to_graph_oracle import lightpath_graph_ref, lightpath_data_ref, topological_data_ref  # noqa: F401
