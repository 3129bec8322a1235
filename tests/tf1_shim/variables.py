"""Weight dict of this repo's layer functions  <->  the TF variable names the reference creates.

``flatten`` is what a feed of explicit weights into the reference needs (tests/golden/make_ref_fixtures.py passes it to
the shim as the variable provider; tests/golden/make_tf1_fixtures.py assigns the same values to real TF variables);
the opposite direction is ``tf_gnn_samples_b200.checkpoint.sort_variables``.
"""
from typing import Dict

import numpy as np


def _auto(base: str, i: int) -> str:
    return base if i == 0 else "%s_%d" % (base, i)


def flatten(weights: Dict, prefix: str = "graph_model/gnn_layer_0/", cell_kind: str = "gru") -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    for l, k in enumerate(weights.get("edge_weights", [])):
        out[prefix + "Edge_%d_Weight/kernel:0" % l] = k                         # gnns/rgcn.py:74 etc.
    for l, a in enumerate(weights.get("attention", [])):
        out[prefix + "Edge_%d_Attention_Parameters:0" % l] = a                  # gnns/rgat.py:76
    for l, k in enumerate(weights.get("film_weights", [])):
        out[prefix + "Edge_%d_FiLM_Computations/kernel:0" % l] = k              # gnns/gnn_film.py:78
    for l, ks in enumerate(weights.get("edge_mlps") or []):
        for j, k in enumerate(ks):
            out[prefix + "Edge_%d_MLP/%s/kernel:0" % (l, _auto("dense", j))] = k   # utils/utils.py:109-118
    for j, k in enumerate(weights.get("aggr_mlp") or []):
        out[prefix + "Aggregation_MLP/%s/kernel:0" % _auto("dense", j)] = k     # gnns/rgin.py:78-82
    for l, per_channel in enumerate(weights.get("channel_weights") or []):
        for c, k in enumerate(per_channel):
            out[prefix + "Edge_%d_Channel_%d_Weight_Computation/kernel:0" % (l, c)] = k   # gnns/rgdcn.py:104
    g, b = weights.get("ln_gamma"), weights.get("ln_beta")
    if g is not None:
        gs = g if isinstance(g, (list, tuple)) else [g]
        bs = b if isinstance(b, (list, tuple)) else [b]
        for t, (gg, bb) in enumerate(zip(gs, bs)):
            out[prefix + "%s/gamma:0" % _auto("LayerNorm", t)] = gg
            out[prefix + "%s/beta:0" % _auto("LayerNorm", t)] = bb
    cell = weights.get("cell")
    if cell is not None:
        scope = "gru_cell" if cell_kind.lower() == "gru" else "simple_rnn_cell"
        for key in ("kernel", "recurrent_kernel", "bias"):
            out[prefix + "%s/%s:0" % (scope, key)] = cell[key]
    return out


def provider_from(named: Dict[str, np.ndarray]):
    """Variable provider for tf1_shim.Session: every variable the reference asks for must be in ``named``."""
    used = set()

    def provide(full_name, shape, init):
        if full_name not in named:
            raise KeyError("the reference created variable %s %s which the weight dict does not provide (have: %s)"
                           % (full_name, shape, sorted(named)))
        used.add(full_name)
        return np.asarray(named[full_name])

    provide.used = used
    return provide
