"""TEST INFRASTRUCTURE ONLY -- the forward + SmoothL1 + backward of TopologicalGNN for ONE graph, written
out by hand (no autograd), in exactly the formulation a fused per-graph training kernel evaluates:
factorised NNConv (T_j = h1_j [P_0 .. P_7 | P_b]), per-destination softmax, every
parameter gradient as an explicit sum.  Checked against autograd of TopologicalGNNOracle
(tests/test_oracle_cpu.py::test_hand_written_backward_matches_autograd); round 2 builds the CUDA kernel
against it.  Follows topological_training/models.py:43-64 and
topological_training/train.py:111-116 (SmoothL1Loss, mean over B x 3)."""
import math

import torch


def _lk(v, s=0.01):
    return torch.where(v > 0, v, s * v)


def _dlk(v, s=0.01):
    return torch.where(v > 0, torch.ones_like(v), torch.full_like(v, s))


def graph_fwd_bwd(p, node_ids, src, dst, attr, y, total_outputs):
    """p: state_dict of TopologicalGNN(Oracle) (fp64 tensors); node_ids [n], src / dst [E] graph-local,
    attr [E,4], y [3]; total_outputs = 3 * B (mean reduction of the loss over the batch).
    Returns (out [3], loss contribution, {parameter name: gradient contribution})."""
    H = p["conv1.lin_query.weight"].shape[0]
    n, E = node_ids.shape[0], src.shape[0]
    Wq, bq = p["conv1.lin_query.weight"], p["conv1.lin_query.bias"]
    Wk, bk = p["conv1.lin_key.weight"], p["conv1.lin_key.bias"]
    Wv, bv = p["conv1.lin_value.weight"], p["conv1.lin_value.bias"]
    We = p["conv1.lin_edge.weight"]
    Ws, bs = p["conv1.lin_skip.weight"], p["conv1.lin_skip.bias"]
    W1, b1 = p["conv2.nn.0.weight"], p["conv2.nn.0.bias"]
    W2, b2 = p["conv2.nn.2.weight"], p["conv2.nn.2.bias"]
    Wroot, bias2 = p["conv2.lin.weight"], p["conv2.bias"]
    Wm1, bm1, Wm2, bm2 = p["mlp.0.weight"], p["mlp.0.bias"], p["mlp.3.weight"], p["mlp.3.bias"]
    K8 = W1.shape[0]
    # ---------------- forward
    X = p["node_embeddings.weight"][node_ids]
    Q, K, V, S = X @ Wq.t() + bq, X @ Wk.t() + bk, X @ Wv.t() + bv, X @ Ws.t() + bs
    Ee = attr @ We.t()
    key = K[src] + Ee
    logit = (Q[dst] * key).sum(-1) / math.sqrt(H)
    alpha = torch.zeros(E, dtype=X.dtype)
    for i in range(n):                                   # softmax over the in-edges of i, in edge order
        m = dst == i
        if m.any():
            z = torch.exp(logit[m] - logit[m].max())
            alpha[m] = z / (z.sum() + 1e-16)
    msg = V[src] + Ee
    O1 = torch.zeros(n, H, dtype=X.dtype).index_add_(0, dst, alpha[:, None] * msg) + S
    H1 = _lk(O1)
    hid = torch.relu(attr @ W1.t() + b1)
    hidp = torch.cat([hid, torch.ones(E, 1, dtype=X.dtype)], 1)                    # [E, 9]
    P = torch.cat([W2.view(H, H, K8).permute(0, 2, 1), b2.view(H, 1, H)], 1)        # P[c][k][o]
    T = torch.einsum("jc,cko->jko", H1, P)                                          # [n, 9, H]
    m_e = torch.einsum("ek,eko->eo", hidp, T[src])
    deg = torch.zeros(n, dtype=X.dtype).index_add_(0, dst, torch.ones(E, dtype=X.dtype)).clamp(min=1)
    O2 = torch.zeros(n, H, dtype=X.dtype).index_add_(0, dst, m_e) / deg[:, None] + H1 @ Wroot.t() + bias2
    H2 = _lk(O2)
    pool = H2.mean(0)
    pre1 = pool @ Wm1.t() + bm1
    Z1 = _lk(pre1)
    out = Z1 @ Wm2.t() + bm2
    diff = out - y
    loss = torch.where(diff.abs() < 1, 0.5 * diff * diff, diff.abs() - 0.5).sum() / total_outputs
    # ---------------- backward
    g = {}
    dout = torch.where(diff.abs() < 1, diff, torch.sign(diff)) / total_outputs
    g["mlp.3.weight"], g["mlp.3.bias"] = torch.outer(dout, Z1), dout
    dpre1 = (Wm2.t() @ dout) * _dlk(pre1)
    g["mlp.0.weight"], g["mlp.0.bias"] = torch.outer(dpre1, pool), dpre1
    dO2 = ((Wm1.t() @ dpre1) / n)[None, :] * _dlk(O2)
    g["conv2.bias"] = dO2.sum(0)
    g["conv2.lin.weight"] = dO2.t() @ H1
    dH1 = dO2 @ Wroot
    dm = dO2[dst] / deg[dst][:, None]
    dhidp = torch.einsum("eko,eo->ek", T[src], dm)
    dT = torch.zeros(n, K8 + 1, H, dtype=X.dtype).index_add_(0, src, hidp[:, :, None] * dm[:, None, :])
    dhidpre = dhidp[:, :K8] * (hid > 0)
    g["conv2.nn.0.weight"], g["conv2.nn.0.bias"] = dhidpre.t() @ attr, dhidpre.sum(0)
    dP = torch.einsum("jc,jko->cko", H1, dT)
    g["conv2.nn.2.weight"] = dP[:, :K8, :].permute(0, 2, 1).reshape(H * H, K8)
    g["conv2.nn.2.bias"] = dP[:, K8, :].reshape(H * H)
    dH1 = dH1 + torch.einsum("cko,jko->jc", P, dT)
    dO1 = dH1 * _dlk(O1)
    g["conv1.lin_skip.weight"], g["conv1.lin_skip.bias"] = dO1.t() @ X, dO1.sum(0)
    dX = dO1 @ Ws
    gd = dO1[dst]                                                                   # upstream of each edge's row
    dalpha = (gd * msg).sum(-1)
    dmsg = alpha[:, None] * gd
    tsum = torch.zeros(n, dtype=X.dtype).index_add_(0, dst, alpha * dalpha)
    dlogit = alpha * (dalpha - tsum[dst])
    s = 1.0 / math.sqrt(H)
    dQ = torch.zeros(n, H, dtype=X.dtype).index_add_(0, dst, dlogit[:, None] * key * s)
    dkey = dlogit[:, None] * Q[dst] * s
    dV = torch.zeros(n, H, dtype=X.dtype).index_add_(0, src, dmsg)
    dK = torch.zeros(n, H, dtype=X.dtype).index_add_(0, src, dkey)
    g["conv1.lin_edge.weight"] = (dmsg + dkey).t() @ attr
    g["conv1.lin_query.weight"], g["conv1.lin_query.bias"] = dQ.t() @ X, dQ.sum(0)
    g["conv1.lin_key.weight"], g["conv1.lin_key.bias"] = dK.t() @ X, dK.sum(0)
    g["conv1.lin_value.weight"], g["conv1.lin_value.bias"] = dV.t() @ X, dV.sum(0)
    dX = dX + dQ @ Wq + dK @ Wk + dV @ Wv
    g["node_embeddings.weight"] = torch.zeros_like(p["node_embeddings.weight"]).index_add_(0, node_ids, dX)
    return out, loss, g
