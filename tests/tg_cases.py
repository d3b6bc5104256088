"""Hand-built network-status samples shared by the CPU (reference vs oracle) and GPU (kernel vs oracle) tests of
the graph construction: a single lightpath, two lightpaths that never share a link, two that share a link at distance
just under / exactly at / over the 0.05 threshold, a link carrying one lightpath on two adjacent channels only (skipped
by to_graph.py:285), a shared link with a self loop, float conn ids (int() truncation), an empty sample."""
import numpy as np


def corner_case_samples():
    from gnn_qot_estimation_b200 import synthetic
    F, L, Q = len(synthetic.LP_FEAT), 4, 12
    fi = {k: i for i, k in enumerate(synthetic.LP_FEAT)}
    freqs = 192.2 + 0.025 * np.arange(Q, dtype=np.float64)

    def lp(conn, src, dst, lut=False):
        v = np.zeros(F, dtype=np.float32)
        v[fi["conn_id"]], v[fi["src_id"]], v[fi["dst_id"]] = conn, src, dst
        v[fi["mod_order"]], v[fi["path_len"]], v[fi["num_spans"]] = 16, 100000 + conn, 3
        v[fi["osnr"]], v[fi["snr"]], v[fi["ber"]] = (-1, -1, -1) if lut else (20.5, 15.25, 1e-4)
        return v

    def sample(places):
        d = np.zeros((F, L, Q), dtype=np.float32)
        for vec, link, q in places:
            v = vec.copy()
            v[fi["freq"]] = freqs[q]
            d[:, link, q] = v
        return d

    a, b, c = lp(7, 1, 2, lut=True), lp(3, 2, 5), lp(11.9, 5, 1)          # 11.9 -> conn id 11
    cases = [
        sample([(a, 0, 0)]),                                               # single lightpath
        sample([(a, 0, 0), (b, 1, 0)]),                                    # no shared link
        sample([(a, 0, 0), (b, 0, 1)]),                                    # 0.025 apart: interact
        sample([(a, 0, 0), (b, 0, 2)]),                                    # 0.05 apart: the strict < boundary
        sample([(a, 0, 0), (b, 0, 3)]),                                    # 0.075 apart: no interaction
        sample([(a, 2, 4), (a, 2, 5), (b, 1, 4)]),                         # a alone on link 2 with two channels: skipped
        sample([(a, 2, 4), (a, 2, 5), (b, 2, 7), (c, 2, 6), (c, 0, 0)]),   # shared link: self loop of a, a-c, c-b
        np.zeros((F, L, Q), dtype=np.float32),                             # empty sample
    ]
    return {"data": np.stack(cases), "target": np.tile(np.array([[20.0, 15.0, 1e-3, 1.0]]), (len(cases), 1)),
            "freqs": freqs, "lp_feat": list(synthetic.LP_FEAT), "metric": list(synthetic.METRICS)}
