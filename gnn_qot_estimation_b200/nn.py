"""Layer modules with PyG's names, constructor arguments and parameter names, backed
by the B200 kernels -- what ``from torch_geometric.nn import global_mean_pool,
TransformerConv, NNConv, GATConv, BatchNorm`` provides to the reference
(topological_training/models.py:3, lightpath_training/models.py:3).

Only the configurations the reference instantiates are implemented; anything else
raises (there is no generic fallback).
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import _lib, ops


class TransformerConv(nn.Module):
    """``TransformerConv(in, out, heads=1, concat=True, beta=False, dropout=0.,
    edge_dim=D, bias=True, root_weight=True)``; state_dict keys
    ``lin_key/lin_query/lin_value.{weight,bias}``, ``lin_edge.weight``,
    ``lin_skip.{weight,bias}``."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, beta=False, dropout=0.0,
                 edge_dim=None, bias=True, root_weight=True):
        super().__init__()
        if heads != 1 or beta or dropout != 0.0 or edge_dim != 4 or not bias or not root_weight \
                or in_channels != out_channels:
            raise NotImplementedError("libqot_b200 TransformerConv: heads=1, beta=False, dropout=0, "
                                      "edge_dim=4, in_channels == out_channels only")
        self.in_channels, self.out_channels, self.heads, self.edge_dim = in_channels, out_channels, 1, edge_dim
        self.lin_key = nn.Linear(in_channels, out_channels)
        self.lin_query = nn.Linear(in_channels, out_channels)
        self.lin_value = nn.Linear(in_channels, out_channels)
        self.lin_edge = nn.Linear(edge_dim, out_channels, bias=False)
        self.lin_skip = nn.Linear(in_channels, out_channels)

    @_lib.on_tensor_device
    def forward(self, x, edge_index, edge_attr=None, *, graph=None, node_ids=None, slope: float = 1.0):
        """``graph`` (ops.GraphIndex) lets callers share one CSR between layers;
        ``node_ids`` fuses the embedding lookup (x is then the embedding table);
        ``slope`` fuses a trailing leaky_relu."""
        if graph is None:
            n = x.shape[0] if node_ids is None else node_ids.shape[0]
            graph = ops.GraphIndex(edge_index, n)
        return ops.transformer_conv(
            x, node_ids, graph, edge_attr,
            self.lin_query.weight, self.lin_query.bias, self.lin_key.weight, self.lin_key.bias,
            self.lin_value.weight, self.lin_value.bias, self.lin_edge.weight,
            self.lin_skip.weight, self.lin_skip.bias, slope)


class NNConv(nn.Module):
    """``NNConv(in, out, nn=Seq(Linear(4,8),ReLU,Linear(8,in*out)), aggr='mean',
    root_weight=True, bias=True)``; keys ``nn.0.*``, ``nn.2.*``,
    ``lin.weight``, ``bias``."""

    def __init__(self, in_channels, out_channels, nn, aggr="add", root_weight=True, bias=True):
        super().__init__()
        ok = (aggr == "mean" and root_weight and bias and in_channels == out_channels and len(nn) == 3
              and isinstance(nn[0], torch.nn.Linear) and isinstance(nn[1], torch.nn.ReLU)
              and isinstance(nn[2], torch.nn.Linear) and nn[0].in_features == 4
              and nn[0].out_features == ops.EDGE_HID and nn[2].in_features == ops.EDGE_HID
              and nn[2].out_features == in_channels * out_channels)
        if not ok:
            raise NotImplementedError("libqot_b200 NNConv: aggr='mean', edge MLP Linear(4,8)-ReLU-"
                                      "Linear(8,H*H), in_channels == out_channels only")
        self.in_channels, self.out_channels, self.aggr = in_channels, out_channels, aggr
        self.nn = nn
        self.lin = torch.nn.Linear(in_channels, out_channels, bias=False)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        bound = 1.0 / math.sqrt(in_channels)            # PyG Linear(weight_initializer='uniform')
        torch.nn.init.uniform_(self.lin.weight, -bound, bound)

    @_lib.on_tensor_device
    def forward(self, x, edge_index, edge_attr=None, *, graph=None, slope: float = 1.0):
        if graph is None:
            graph = ops.GraphIndex(edge_index, x.shape[0])
        return ops.nnconv_mean(x, graph, edge_attr, self.nn[0].weight, self.nn[0].bias,
                               self.nn[2].weight, self.nn[2].bias, self.lin.weight, self.bias, slope)


class GATConv(nn.Module):
    """``GATConv(in, out, heads=4, concat=True, negative_slope=0.2, dropout=0.,
    add_self_loops=True, bias=True)``; keys ``lin.weight``,
    ``att_src``, ``att_dst`` ([1,H,C]), ``bias``."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, add_self_loops=True, bias=True):
        super().__init__()
        if not concat or negative_slope != 0.2 or dropout != 0.0 or not add_self_loops or not bias:
            raise NotImplementedError("libqot_b200 GATConv: concat=True, negative_slope=0.2, dropout=0, "
                                      "add_self_loops=True, bias=True only")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels))
        nn.init.xavier_uniform_(self.lin.weight)         # PyG glorot
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)

    @_lib.on_tensor_device
    def forward(self, x, edge_index, *, graph=None):
        if (self.in_channels, self.out_channels, self.heads) != (5, 32, 4):
            raise NotImplementedError("libqot_b200 GATConv kernels: in=5, out=32, heads=4 only")
        if graph is None:
            graph = ops.GraphIndex(edge_index, x.shape[0])
        return ops.gat_conv(x, graph, self.lin.weight, self.att_src, self.att_dst, self.bias)


class BatchNorm(nn.Module):
    """PyG ``BatchNorm(C)``: wraps ``torch.nn.BatchNorm1d`` as ``.module`` (keys
    ``module.weight/bias/running_mean/running_var/num_batches_tracked``)."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        if not affine or not track_running_stats:
            raise NotImplementedError("libqot_b200 BatchNorm: affine=True, track_running_stats=True only")
        self.module = nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    @_lib.on_tensor_device
    def forward(self, x):
        return ops.batch_norm(x, self.module, self.training)


@_lib.on_tensor_device
def global_mean_pool(x, batch, size=None):
    """Per-graph mean of node rows."""
    B = int(size) if size is not None else (int(batch.max().item()) + 1 if batch.numel() else 0)
    return ops.mean_pool(x, ops.graph_ptr(batch, B))
