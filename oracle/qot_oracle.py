"""CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py): a pure-PyTorch
restatement of the reference's message-passing hot path, using the same op
decomposition PyTorch Geometric uses (index_select gathers, scatter_add /
scatter_reduce(amax), an [E,H,H] bmm for NNConv).

PARITY UNPINNED (PyG absent; no golden vectors in the reference) -- see the
package docstring.  Every function cites the reference call site it follows and
the SURVEY.md Appendix-A paragraph holding the PyG semantics it restates.

All functions are dtype-agnostic (fp32 for parity/timing, fp64 for error
budgeting and gradcheck) and differentiable through torch autograd, so the
oracle also provides reference gradients for the backward kernels.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = [
    "scatter_sum", "scatter_amax", "segment_softmax",
    "collate_ref", "build_csr_ref", "gat_edges_ref", "graph_ptr_ref",
    "transformer_conv_ref", "nnconv_mean_ref", "nnconv_mean_factorised_ref",
    "gat_conv_ref", "global_mean_pool_ref", "batchnorm_ref", "lut_select_ref",
    "OracleTransformerConv", "OracleNNConv", "OracleGATConv", "OracleBatchNorm",
    "TopologicalGNNOracle", "LightpathGNNOracle",
]


# --------------------------------------------------------------------------- #
# scatter primitives (PyG utils.scatter / utils.softmax)
# --------------------------------------------------------------------------- #
def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """``scatter(src, index, 0, dim_size, 'sum')``: rows receiving nothing are 0."""
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def scatter_amax(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """``scatter(src, index, 0, dim_size, 'max')``; untouched rows are 0 like PyG."""
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    return out.scatter_reduce(0, idx, src, reduce="amax", include_self=False)


def segment_softmax(src: torch.Tensor, index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """PyG ``utils.softmax(src, index, num_nodes=N)`` (Appendix A.0):
    m = scatter_max(src.detach()); z = exp(src - m[index]);
    s = scatter_sum(z) + 1e-16; z / s[index]."""
    m = scatter_amax(src.detach(), index, num_nodes)
    z = (src - m.index_select(0, index)).exp()
    s = scatter_sum(z, index, num_nodes) + 1e-16
    return z / s.index_select(0, index)


# --------------------------------------------------------------------------- #
# collate / CSR (integer work: numpy; bit-exact target for the CUDA collate)
# --------------------------------------------------------------------------- #
def collate_ref(graphs: Sequence[dict]) -> SimpleNamespace:
    """PyG ``Batch.from_data_list`` as driven by the reference DataLoaders
    (topological_training/train.py:93-95, lightpath_training/train.py:94-96;
    Appendix A.6).  ``graphs`` is a list of dicts holding the per-graph Data
    fields the reference datasets emit (topological_training/dataset.py:75-123,
    lightpath_training/dataset.py:86-123): ``edge_index [2,E_g]`` int64 and any
    of ``x [n,F]``, ``edge_attr [E_g,D]``, ``node_ids [n]``, ``y``; ``num_nodes``.

    Keys containing 'index' are concatenated on the last dim and offset by the
    running node count; everything else is concatenated on dim 0 unchanged.
    ``batch`` = repeat_interleave(arange(B), n_g); ``ptr`` = node offsets.
    """
    B = len(graphs)
    n = [int(g["num_nodes"]) for g in graphs]
    ptr = np.zeros(B + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(n)
    out = SimpleNamespace()
    out.num_graphs = B
    out.ptr = torch.from_numpy(ptr.copy())
    out.batch = torch.repeat_interleave(torch.arange(B, dtype=torch.int64),
                                        torch.tensor(n, dtype=torch.int64))
    out.edge_index = torch.cat(
        [g["edge_index"].to(torch.int64) + int(ptr[i]) for i, g in enumerate(graphs)], dim=1
    ) if B else torch.zeros(2, 0, dtype=torch.int64)
    eptr = np.zeros(B + 1, dtype=np.int64)
    eptr[1:] = np.cumsum([int(g["edge_index"].shape[1]) for g in graphs])
    out.edge_ptr = torch.from_numpy(eptr)
    for key in ("x", "edge_attr", "node_ids", "y"):
        vals = [g.get(key) for g in graphs]
        if B and all(v is not None for v in vals):
            setattr(out, key, torch.cat(list(vals), dim=0))
        else:
            setattr(out, key, None)
    return out


def graph_ptr_ref(batch: torch.Tensor, num_graphs: int) -> torch.Tensor:
    """Node offsets per graph from a sorted ``batch`` vector (PyG ``Batch.ptr``)."""
    counts = torch.bincount(batch, minlength=num_graphs)
    ptr = torch.zeros(num_graphs + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr


def build_csr_ref(edge_index: torch.Tensor, num_nodes: int):
    """Destination-sorted CSR of ``edge_index`` (row = target i, entries = in-edges
    j->i in ORIGINAL edge order, i.e. a stable sort by destination).  This is the
    order in which PyG's scatter visits a destination's messages sequentially on
    CPU.  Returns int32 ``rowptr [N+1]``, ``src [E]``, ``eid [E]``."""
    src = edge_index[0].cpu().numpy().astype(np.int64)
    dst = edge_index[1].cpu().numpy().astype(np.int64)
    order = np.argsort(dst, kind="stable")
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    return (torch.from_numpy(rowptr.astype(np.int32)),
            torch.from_numpy(src[order].astype(np.int32)),
            torch.from_numpy(order.astype(np.int32)))


def gat_edges_ref(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """GATConv's edge preprocessing (Appendix A.3): ``remove_self_loops`` (mask
    src != dst, order preserved) then ``add_self_loops`` (append [0..N-1] twice)."""
    keep = edge_index[0] != edge_index[1]
    ei = edge_index[:, keep]
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([ei, torch.stack([loops, loops])], dim=1)


# --------------------------------------------------------------------------- #
# layers (functional)
# --------------------------------------------------------------------------- #
def transformer_conv_ref(x, edge_index, edge_attr, Wq, bq, Wk, bk, Wv, bv, We, Ws, bs):
    """PyG ``TransformerConv(H, H, heads=1, concat=True, beta=False, edge_dim=4,
    root_weight=True)`` as called at topological_training/models.py:15-17,53
    (Appendix A.1)."""
    N, C = x.shape[0], Wq.shape[0]
    src, dst = edge_index[0], edge_index[1]
    q = F.linear(x, Wq, bq)
    k = F.linear(x, Wk, bk)
    v = F.linear(x, Wv, bv)
    e = F.linear(edge_attr, We)                      # lin_edge has no bias
    q_i = q.index_select(0, dst)
    k_j = k.index_select(0, src) + e
    alpha = (q_i * k_j).sum(-1) / math.sqrt(C)
    alpha = segment_softmax(alpha, dst, N)
    msg = (v.index_select(0, src) + e) * alpha.unsqueeze(-1)
    out = scatter_sum(msg, dst, N)
    return out + F.linear(x, Ws, bs)                 # lin_skip(x)


def nnconv_mean_ref(x, edge_index, edge_attr, W1, b1, W2, b2, Wroot, bias):
    """PyG ``NNConv(H, H, nn=Seq(Linear(4,8),ReLU,Linear(8,H*H)), aggr='mean')`` as
    called at topological_training/models.py:20-30,57 (Appendix A.2).  Direct
    form: materialises the per-edge [E,H_in,H_out] weight like PyG does."""
    N, Hin = x.shape
    Hout = Wroot.shape[0]
    src, dst = edge_index[0], edge_index[1]
    h = F.relu(F.linear(edge_attr, W1, b1))
    W = F.linear(h, W2, b2).view(-1, Hin, Hout)      # flat index i_in*H_out + o
    msg = torch.matmul(x.index_select(0, src).unsqueeze(1), W).squeeze(1)
    summed = scatter_sum(msg, dst, N)
    cnt = scatter_sum(torch.ones_like(dst, dtype=x.dtype), dst, N).clamp(min=1)
    out = summed / cnt.unsqueeze(-1)
    return out + F.linear(x, Wroot) + bias


def nnconv_mean_factorised_ref(x, edge_index, edge_attr, W1, b1, W2, b2, Wroot, bias,
                               chunk: int = 65536):
    """Same result as :func:`nnconv_mean_ref` without the [E,H,H] temporary
    (Appendix A.2 'factorised identity'): msg_e = sum_k h_e[k]*(x_j P_k) + x_j P_b
    with P_k[i,o] = W2[i*H+o,k], P_b[i,o] = b2[i*H+o].  Used where the direct
    form is infeasible (BASELINE cfg 5: 21 GB) and cross-checked against it at
    small sizes."""
    N, Hin = x.shape
    Hout = Wroot.shape[0]
    K = W1.shape[0]
    src, dst = edge_index[0], edge_index[1]
    h = F.relu(F.linear(edge_attr, W1, b1))                            # [E,K]
    P = W2.view(Hin, Hout, K).permute(2, 0, 1)                        # [K,Hin,Hout]
    Pb = b2.view(Hin, Hout)
    summed = x.new_zeros(N, Hout)
    E = src.numel()
    # node-wise products y[n,k,:] = x[n] @ P_k  (+ bias term as k = K)
    Y = torch.einsum("ni,kio->nko", x, torch.cat([P, Pb.unsqueeze(0)], 0))  # [N,K+1,Hout]
    for s in range(0, E, chunk):
        sl = slice(s, min(E, s + chunk))
        hj = torch.cat([h[sl], h.new_ones(h[sl].shape[0], 1)], 1)      # [e,K+1]
        msg = torch.einsum("ek,eko->eo", hj, Y.index_select(0, src[sl]))
        summed.index_add_(0, dst[sl], msg)
    cnt = scatter_sum(torch.ones_like(dst, dtype=x.dtype), dst, N).clamp(min=1)
    return summed / cnt.unsqueeze(-1) + F.linear(x, Wroot) + bias


def gat_conv_ref(x, edge_index, W, att_src, att_dst, bias, negative_slope: float = 0.2):
    """PyG ``GATConv(5, 32, heads=4, concat=True)`` as called at
    lightpath_training/models.py:13,30 (Appendix A.3).  ``att_*`` are [1,H,C]."""
    N = x.shape[0]
    Hh, C = att_src.shape[-2], att_src.shape[-1]
    xp = F.linear(x, W).view(N, Hh, C)                                  # lin has no bias
    a_src = (xp * att_src.view(1, Hh, C)).sum(-1)                       # [N,H]
    a_dst = (xp * att_dst.view(1, Hh, C)).sum(-1)
    ei = gat_edges_ref(edge_index, N)
    src, dst = ei[0], ei[1]
    alpha = a_src.index_select(0, src) + a_dst.index_select(0, dst)     # [E',H]
    alpha = F.leaky_relu(alpha, negative_slope)
    alpha = segment_softmax(alpha, dst, N)
    msg = xp.index_select(0, src) * alpha.unsqueeze(-1)                 # [E',H,C]
    out = scatter_sum(msg, dst, N).view(N, Hh * C)
    return out + bias


def global_mean_pool_ref(x, batch, num_graphs: Optional[int] = None):
    """PyG ``global_mean_pool`` at topological_training/models.py:61 (Appendix A.4):
    B = batch.max()+1 unless given; mean with count clamped to >= 1."""
    B = int(batch.max()) + 1 if num_graphs is None else int(num_graphs)
    s = scatter_sum(x, batch, B)
    cnt = scatter_sum(torch.ones_like(batch, dtype=x.dtype), batch, B).clamp(min=1)
    return s / cnt.unsqueeze(-1)


def batchnorm_ref(x, weight, bias, running_mean, running_var, training: bool,
                  momentum: float = 0.1, eps: float = 1e-5):
    """PyG ``BatchNorm(128)`` == ``torch.nn.BatchNorm1d`` over the node dimension
    (lightpath_training/models.py:14,31; Appendix A.5).  Updates the running
    buffers in place when training."""
    return F.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)


def lut_select_ref(x_feat, h, batch, is_lut_index: int):
    """LUT readout at lightpath_training/models.py:35-40."""
    lut_mask = x_feat[:, is_lut_index] == 1.0
    if not bool(lut_mask.any()):
        raise ValueError("No LUT node found in the batch.")
    return h[lut_mask], batch[lut_mask]


# --------------------------------------------------------------------------- #
# nn.Module restatements with the shipped checkpoints' parameter names
# --------------------------------------------------------------------------- #
class OracleTransformerConv(nn.Module):
    def __init__(self, in_channels, out_channels, edge_dim):
        super().__init__()
        self.lin_key = nn.Linear(in_channels, out_channels)
        self.lin_query = nn.Linear(in_channels, out_channels)
        self.lin_value = nn.Linear(in_channels, out_channels)
        self.lin_edge = nn.Linear(edge_dim, out_channels, bias=False)
        self.lin_skip = nn.Linear(in_channels, out_channels)

    def forward(self, x, edge_index, edge_attr):
        return transformer_conv_ref(
            x, edge_index, edge_attr,
            self.lin_query.weight, self.lin_query.bias,
            self.lin_key.weight, self.lin_key.bias,
            self.lin_value.weight, self.lin_value.bias,
            self.lin_edge.weight, self.lin_skip.weight, self.lin_skip.bias)


class OracleNNConv(nn.Module):
    def __init__(self, in_channels, out_channels, edge_nn, factorised=False):
        super().__init__()
        self.nn = edge_nn
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels))
        bound = 1.0 / math.sqrt(in_channels)          # PyG Linear 'uniform' initializer
        nn.init.uniform_(self.lin.weight, -bound, bound)
        self.factorised = factorised

    def forward(self, x, edge_index, edge_attr):
        fn = nnconv_mean_factorised_ref if self.factorised else nnconv_mean_ref
        return fn(x, edge_index, edge_attr,
                  self.nn[0].weight, self.nn[0].bias, self.nn[2].weight, self.nn[2].bias,
                  self.lin.weight, self.bias)


class OracleGATConv(nn.Module):
    def __init__(self, in_channels, out_channels, heads):
        super().__init__()
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels))
        nn.init.xavier_uniform_(self.lin.weight)       # PyG glorot
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)

    def forward(self, x, edge_index):
        return gat_conv_ref(x, edge_index, self.lin.weight, self.att_src, self.att_dst, self.bias)


class OracleBatchNorm(nn.Module):
    """PyG BatchNorm keeps the torch BatchNorm1d as ``.module`` (checkpoint keys
    ``norm1.module.*``)."""
    def __init__(self, channels):
        super().__init__()
        self.module = nn.BatchNorm1d(channels)

    def forward(self, x):
        return self.module(x)


class TopologicalGNNOracle(nn.Module):
    """Restates topological_training/models.py:6-64."""
    def __init__(self, num_nodes, hidden_channels, out_channels, edge_dim, dropout_p=0.5,
                 factorised_nnconv=False):
        super().__init__()
        self.node_embeddings = nn.Embedding(num_nodes, hidden_channels)
        self.conv1 = OracleTransformerConv(hidden_channels, hidden_channels, edge_dim)
        edge_nn = nn.Sequential(
            nn.Linear(edge_dim, edge_dim * 2), nn.ReLU(),
            nn.Linear(edge_dim * 2, hidden_channels * hidden_channels))
        self.conv2 = OracleNNConv(hidden_channels, hidden_channels, edge_nn, factorised_nnconv)
        self.mlp = nn.Sequential(
            nn.Linear(hidden_channels, hidden_channels), nn.LeakyReLU(),
            nn.Dropout(p=dropout_p), nn.Linear(hidden_channels, out_channels))
        self.dropout = nn.Dropout(p=dropout_p)

    def forward(self, data, masks=None):
        """`masks` (test hook): (m1 [N,H], m2 [N,H], m3 [B,H], scale, scale_head) -- explicit dropout masks (1 = keep)
        in place of the three nn.Dropout draws of models.py:55,59 and :36-41, so that a kernel's masks can be
        replayed here; dropout(x) = x * mask / (1 - p)."""
        x, edge_index, edge_attr, batch = data.x, data.edge_index, data.edge_attr, data.batch
        if x is None or x.numel() == 0:
            x = self.node_embeddings(data.node_ids)
        if masks is None:
            x = self.dropout(F.leaky_relu(self.conv1(x, edge_index, edge_attr)))
            x = self.dropout(F.leaky_relu(self.conv2(x, edge_index, edge_attr)))
            x = global_mean_pool_ref(x, batch)
            return self.mlp(x)
        m1, m2, m3, scale, scale_h = masks
        x = F.leaky_relu(self.conv1(x, edge_index, edge_attr)) * (m1.to(x.dtype) * scale)
        x = F.leaky_relu(self.conv2(x, edge_index, edge_attr)) * (m2.to(x.dtype) * scale)
        x = global_mean_pool_ref(x, batch)
        h = self.mlp[1](self.mlp[0](x)) * (m3.to(x.dtype) * scale_h)
        return self.mlp[3](h)


class LightpathGNNOracle(nn.Module):
    """Restates lightpath_training/models.py:7-45."""
    def __init__(self, in_channels, hidden_channels, output_dim, is_lut_index, dropout_p=0.5):
        super().__init__()
        self.conv1 = OracleGATConv(in_channels, hidden_channels, heads=4)
        self.norm1 = OracleBatchNorm(hidden_channels * 4)
        self.mlp = nn.Sequential(
            nn.Linear(hidden_channels * 4, hidden_channels), nn.LeakyReLU(),
            nn.Dropout(p=dropout_p), nn.Linear(hidden_channels, output_dim))
        self.is_lut_index = is_lut_index

    def forward(self, data):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        h = F.relu(self.norm1(self.conv1(x, edge_index)))
        lut_embedding, lut_batch = lut_select_ref(data.x, h, batch, self.is_lut_index)
        return self.mlp(lut_embedding), lut_batch
