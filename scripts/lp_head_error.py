"""Max abs / relative error of the LightpathGNN eval kernel's two readout heads (tcgen05 in-kernel head, FP32 split
head) against the fp64 oracle on 8192 synthetic graphs with the shipped weights.  GPU only (oracle on CPU)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_golden  # noqa: E402
from gnn_qot_estimation_b200 import LightpathGNN, synthetic  # noqa: E402
from oracle import LightpathGNNOracle  # noqa: E402

dev = torch.device("cuda:0")
sd = load_golden("ckpt_lightpath_model_1.pt")["model_state_dict"]
m = LightpathGNN(5, 32, 3, is_lut_index=1, dropout_p=0.0); m.load_state_dict(sd); m.to(dev).eval()
o = LightpathGNNOracle(5, 32, 3, is_lut_index=1, dropout_p=0.0).double(); o.load_state_dict(sd); o.eval()
G = 8192
store = synthetic.lightpath_store(G, seed=3, device="cpu", lut_per_graph=1)
hb = store.host_batch(0, G); hb.x = hb.x.double()
with torch.no_grad():
    eo, _ = o(hb)
db = store.to(dev).collate(range(G))
for split in (False, True):
    plan = m.stream_plan([db], split_head=split)
    m.forward_stream(plan); torch.cuda.synchronize()
    out = plan.result(0).out[:G].cpu().double()
    d = (out - eo).abs()
    print(f"{'fp32 split head' if split else 'tcgen05 head   '}: max abs err {d.max():.3e}  max rel-to-max {d.max() / eo.abs().max():.3e}  "
          f"mean abs {d.mean():.3e}  |out| max {eo.abs().max():.3f}")
