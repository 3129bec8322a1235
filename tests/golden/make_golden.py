"""Regenerates tests/golden/*.npz -- small seeded input/output vectors for every layer.

The reference cannot run in this image (TF1 / dpu_utils missing: SURVEY.md 0), so these vectors come from
the float64 oracle (oracle/ref_layers.py), not from the reference itself; they pin the oracle against
accidental change and let the GPU tests compare against committed numbers.  Usage:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_layers as R          # noqa: E402
from tf_gnn_samples_b200 import weights as W  # noqa: E402
from helpers import node_states, tiny_graph   # noqa: E402

V, D, EDGES = 48, 32, (110, 48, 0, 70)


def flatten(prefix, obj, out):
    if isinstance(obj, dict):
        for k, v in obj.items():
            flatten("%s.%s" % (prefix, k), v, out)
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            flatten("%s.%d" % (prefix, i), v, out)
    else:
        out[prefix] = np.asarray(obj)


CASES = {
    "rgcn": dict(fn="rgcn", wfn=lambda: W.rgcn_weights(4, D, D), indeg=True,
                 kw=dict(state_dim=D, num_timesteps=2, activation_function="tanh")),
    "ggnn": dict(fn="ggnn", wfn=lambda: W.ggnn_weights(4, D, random_bias=True), indeg=False,
                 kw=dict(state_dim=D, num_timesteps=3, gated_unit_type="gru", activation_function="tanh")),
    "rgat": dict(fn="rgat", wfn=lambda: W.rgat_weights(4, D, D), indeg=False,
                 kw=dict(state_dim=D, num_heads=4, activation_function="tanh")),
    "gnn-film": dict(fn="gnn-film", wfn=lambda: W.film_weights(4, D, D, random_ln=True), indeg=True,
                     kw=dict(state_dim=D, activation_function="ReLU", normalize_by_num_incoming=True)),
    "gnn-edge-mlp": dict(fn="gnn-edge-mlp", wfn=lambda: W.edge_mlp_weights(4, D, D, random_ln=True), indeg=True,
                         kw=dict(state_dim=D, activation_function="gelu", num_edge_hidden_layers=1)),
    "rgin": dict(fn="rgin", wfn=lambda: W.rgin_weights(4, D, D, num_aggr_MLP_hidden_layers=1, random_ln=True), indeg=False,
                 kw=dict(state_dim=D, activation_function="ReLU", num_edge_MLP_hidden_layers=1, num_aggr_MLP_hidden_layers=1)),
    "rgdcn": dict(fn="rgdcn", wfn=lambda: W.rgdcn_weights(4, 4, 8, stddev=0.15), indeg=True,
                  kw=dict(num_channels=4, channel_dim=8, num_timesteps=2, activation_function="tanh")),
}


def build_case(name):
    c = CASES[name]
    adj, indeg = tiny_graph(V, EDGES, seed=31)
    h = node_states(V, D, seed=32)
    w = c["wfn"]()
    args = (indeg,) if c["indeg"] else ()
    out = R.LAYERS[c["fn"]](h, adj, *args, **c["kw"], weights=w)
    return h, adj, indeg, w, c["kw"], c["indeg"], out


def main():
    for name in CASES:
        h, adj, indeg, w, kw, use_indeg, out = build_case(name)
        blob = {"h": h, "indeg": indeg, "out": out.astype(np.float64)}
        for l, a in enumerate(adj):
            blob["adj.%d" % l] = a
        flatten("w", w, blob)
        np.savez_compressed(os.path.join(HERE, "%s.npz" % name.replace("-", "_")), **blob)
        print(name, out.shape, float(np.abs(out).max()))


if __name__ == "__main__":
    main()
