"""LightpathGNN on the B200 kernels -- drop-in for lightpath_training/models.py:7-45.

Same constructor, ``forward(data) -> (out [L,3], lut_batch [L])``, same
``ValueError("No LUT node found in the batch.")`` and the same state_dict names as
the shipped checkpoints (``conv1.lin.weight``, ``conv1.att_src``, ``conv1.att_dst``,
``conv1.bias``, ``norm1.module.*``, ``mlp.0.*``, ``mlp.3.*``), so
``lightpath_training/test.py:63`` loads ``models/model_N.pth`` with strict=True.

eval(): one launch of the persistent kernel (csrc/lightpath_stream.cu) per batch -- or per MANY
batches through ``stream_plan`` / ``forward_stream``.
train(): GAT forward -> batch statistics -> LUT head, with the matching backward
kernels (csrc/lightpath_train.cu) through ``torch.autograd.Function``.
"""
from __future__ import annotations

import torch
from torch.nn import Dropout, LeakyReLU, Linear

from .. import _lib, ops
from ..nn import BatchNorm, GATConv


class LightpathGNN(torch.nn.Module):
    dominant_kernel = "lp_stream_kernel"   # csrc/lightpath_stream.cu: eval forward, one launch for any number of batches

    def __init__(self, in_channels, hidden_channels, output_dim, is_lut_index, dropout_p=0.5):
        super().__init__()
        self.conv1 = GATConv(in_channels, hidden_channels, heads=4, concat=True)
        self.norm1 = BatchNorm(hidden_channels * 4)
        self.mlp = torch.nn.Sequential(
            Linear(hidden_channels * 4, hidden_channels),
            LeakyReLU(),
            Dropout(p=dropout_p),
            Linear(hidden_channels, output_dim),
        )
        self.is_lut_index = is_lut_index
        self._prepared = None
        self._prepared_key = None

    # ------------------------------------------------------------------ eval
    def _eval_params(self):
        bn = self.norm1.module
        return {
            "lin_w": self.conv1.lin.weight, "att_src": self.conv1.att_src, "att_dst": self.conv1.att_dst,
            "conv_bias": self.conv1.bias, "bn_w": bn.weight, "bn_b": bn.bias,
            "bn_mean": bn.running_mean, "bn_var": bn.running_var,
            "mlp_w1": self.mlp[0].weight, "mlp_b1": self.mlp[0].bias,
            "mlp_w2": self.mlp[3].weight, "mlp_b2": self.mlp[3].bias,
        }

    def prepared(self) -> torch.Tensor:
        """Folded eval-mode parameters, rebuilt only when a tensor changed."""
        ps = self._eval_params()
        key = tuple((t.data_ptr(), t._version) for t in ps.values())
        if self._prepared is None or key != self._prepared_key:
            self._prepared = ops.lightpath_prepare(ps, self.norm1.module.eps, self.is_lut_index)
            self._prepared_key = key
        return self._prepared

    @_lib.on_tensor_device
    def forward_device(self, data) -> ops.LightpathInferOut:
        """Eval forward of ONE batch without reading anything back: one launch of the persistent kernel
        (qot_lightpath_infer_stream, n_batches = 1) when the batch carries ``ptr`` / ``edge_ptr`` / ``lut_ptr``
        (every collate of this package provides them; a foreign batch gets them from qot_graph_ptr / qot_edge_ptr /
        qot_lightpath_lut_ptr first).  The plan (output buffers + device descriptor) is cached on the batch object."""
        from ..batch import Batch
        cache = ops.batch_cache(data)
        key = ("stream_plan", self.is_lut_index, data.x.data_ptr(), data.edge_index.data_ptr(),
               tuple(data.x.shape), tuple(data.edge_index.shape))
        plan = cache.get(key)
        if plan is None:
            gptr = ops.batch_graph_ptr(data)
            eptr = getattr(data, "edge_ptr", None)
            if eptr is None:
                if "eptr" not in cache:
                    cache["eptr"] = ops.edge_ptr(data.edge_index, data.batch, gptr.numel() - 1)
                eptr = cache["eptr"][0]
            lut_ptr = getattr(data, "lut_ptr", None)
            if lut_ptr is None or getattr(data, "lut_col", None) != self.is_lut_index:
                lut_ptr = ops.lightpath_lut_ptr(data.x, gptr, self.is_lut_index)
            view = Batch(x=data.x, edge_index=data.edge_index, ptr=gptr, edge_ptr=eptr, lut_ptr=lut_ptr,
                         num_graphs=int(gptr.numel() - 1), lut_col=self.is_lut_index,
                         sym_by_src=bool(getattr(data, "sym_by_src", False)))
            plan = cache[key] = ops.LightpathStreamPlan([view], self.is_lut_index)
        plan.launch(self.prepared())
        return plan.result(0)

    def stream_plan(self, batches, split_head: bool = False) -> "ops.LightpathStreamPlan":
        """Plan for evaluating many resident batches with one launch of the persistent kernel each time
        ``forward_stream`` is called (the streaming form of lightpath_training/test.py:77-94)."""
        self._check_supported()
        return ops.LightpathStreamPlan(batches, self.is_lut_index, split_head)

    @_lib.on_tensor_device
    def forward_stream(self, plan, first: int = 0, count=None) -> None:
        """Enqueues batches [first, first+count) of ``plan`` (eval mode; no host synchronisation)."""
        if self.training:
            raise RuntimeError("LightpathGNN.forward_stream is the eval-mode forward: call model.eval()")
        plan.launch(self.prepared(), first, count)

    def _forward_eval(self, data):
        self._check_supported()
        res = self.forward_device(data)
        flags = [res.n_lut, res.status]
        if getattr(data, "edge_ptr", None) is None:
            flags.append(ops.batch_cache(data)["eptr"][1])      # "edges grouped by graph" check
        vals = torch.cat(flags).tolist()                          # the one D2H of this forward
        n_lut, stale = vals[0], vals[1]
        if len(vals) > 2 and vals[2]:
            return self._forward_general(data)                    # CSR path handles any edge order
        if stale:
            raise RuntimeError("LightpathGNN: the batch's lut_ptr does not match x[:, is_lut_index] "
                               "(stale or foreign index array)")
        if n_lut == 0:
            raise ValueError("No LUT node found in the batch.")
        return res.out[:n_lut].clone(), res.lut_batch[:n_lut].clone()   # the plan's buffers are reused by the next call

    def _check_supported(self):
        c = self.conv1
        if (c.in_channels, c.out_channels, c.heads) != (5, 32, 4) or self.mlp[0].out_features != 32 \
                or self.mlp[3].out_features != 3:
            raise RuntimeError("libqot_b200 implements the reference LightpathGNN shape only: "
                               "in_channels=5, hidden_channels=32, heads=4, output_dim=3")

    # ----------------------------------------------------------- train/general
    def _forward_general(self, data):
        self._check_supported()
        h = self.conv1(data.x, data.edge_index)
        out, lut_batch = ops.lut_bn_head(
            h, data.x, data.batch, self.is_lut_index, self.norm1.module,
            self.mlp[0].weight, self.mlp[0].bias, self.mlp[3].weight, self.mlp[3].bias,
            self.training, self.mlp[2].p if self.training else 0.0,
            getattr(data, "lut_rows", None) if getattr(data, "lut_col", None) == self.is_lut_index else None)
        return out, lut_batch

    @_lib.on_tensor_device
    def forward(self, data):
        if not data.x.is_cuda:
            raise RuntimeError("LightpathGNN (B200) needs the batch on a CUDA device; there is no CPU path")
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if self.training or needs_grad:
            return self._forward_general(data)
        return self._forward_eval(data)
