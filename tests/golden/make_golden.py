"""Generates the committed fixtures under tests/golden/ (run in the BUILD container,
where /root/reference is mounted; the GPU box never reads /root/reference).

  ckpt_*.pt            the three shipped checkpoints' state_dicts + model_params,
                       re-saved as plain tensor dicts (weights are data, not source)
  lightpath_eval.pt    seeded lightpath batch + oracle outputs with model_1 weights
  topological_train.pt seeded NSFNET batch + oracle out / loss / every gradient with
                       model_0 weights (fp32 oracle and fp64 oracle)

The reference's own implementation (PyTorch Geometric) cannot be imported here, so the
expected values come from oracle/ (PARITY UNPINNED at the PyG boundary -- see
oracle/__init__.py); the checkpoints pin names, shapes and realistic weight values.
"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import LightpathGNNOracle, TopologicalGNNOracle  # noqa: E402
from gnn_qot_estimation_b200 import synthetic  # noqa: E402

REF = Path(os.environ.get("QOT_REFERENCE", "/root/reference"))
OUT = Path(__file__).resolve().parent

CKPTS = {
    "ckpt_topological_model_0.pt": REF / "topological_training/models/model_0.pth",
    "ckpt_lightpath_model_0.pt": REF / "lightpath_training/models/model_0.pth",
    "ckpt_lightpath_model_1.pt": REF / "lightpath_training/models/model_1.pth",
}


def batch_dict(b):
    return {k: getattr(b, k) for k in ("x", "edge_index", "edge_attr", "batch", "node_ids", "y", "ptr", "edge_ptr")
            if getattr(b, k) is not None} | {"num_graphs": b.num_graphs}


def main():
    torch.manual_seed(0)
    for name, src in CKPTS.items():
        ck = torch.load(src, map_location="cpu", weights_only=False)
        torch.save({"model_state_dict": {k: v.clone() for k, v in ck["model_state_dict"].items()},
                    "model_params": ck["model_params"]}, OUT / name)

    # ---- lightpath eval (cfg-2 generator, small) ----
    ck = torch.load(OUT / "ckpt_lightpath_model_1.pt")
    store = synthetic.lightpath_store(96, seed=1, device="cpu")
    b = store.host_batch(0, 96)
    res = {}
    for dt in (torch.float32, torch.float64):
        m = LightpathGNNOracle(5, 32, 3, is_lut_index=1, dropout_p=0.0).to(dt)
        m.load_state_dict(ck["model_state_dict"], strict=True)
        m.eval()
        bb = b.to("cpu")
        bb.x = bb.x.to(dt)
        with torch.no_grad():
            out, lut_batch = m(bb)
        res[str(dt)] = {"out": out, "lut_batch": lut_batch}
    torch.save({"batch": batch_dict(b), "expected": res}, OUT / "lightpath_eval.pt")

    # ---- topological fwd + SmoothL1 + bwd (cfg-1 graphs, shipped weights) ----
    ck = torch.load(OUT / "ckpt_topological_model_0.pt")
    store = synthetic.nsfnet_store(64, seed=0, device="cpu")
    b = store.host_batch(0, 64)
    res = {}
    for dt in (torch.float32, torch.float64):
        m = TopologicalGNNOracle(75, 16, 3, edge_dim=4, dropout_p=0.0).to(dt)
        m.load_state_dict(ck["model_state_dict"], strict=True)
        m.train()
        bb = b.to("cpu")
        bb.edge_attr = bb.edge_attr.to(dt)
        out = m(bb)
        loss = torch.nn.SmoothL1Loss()(out, bb.y.to(dt).view(-1, 3))
        loss.backward()
        res[str(dt)] = {"out": out.detach(), "loss": loss.detach(),
                        "grads": {k: p.grad.clone() for k, p in m.named_parameters()}}
    torch.save({"batch": batch_dict(b), "expected": res}, OUT / "topological_train.pt")
    for p in sorted(OUT.glob("*.pt")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()
