"""Reference checkpoint compatibility.

The reference snapshots a model as a pickle (models/sparse_graph_model.py:91-108):
    {"model_class", "task_class", "model_params", "task_params", "task_metadata",
     "weights": {tf variable name -> numpy array}}
and restores it by variable NAME (``load_weights``, :110-126; utils/model_utils.py:58-77).  Kernels are Keras
``[in, out]`` matrices -- the orientation this engine uses -- so nothing is transposed.

This module reads such a pickle without TensorFlow, sorts the variables into the ``weights=`` dictionaries the
layer functions of ``gnns/`` take, and writes the same structure back.  Variable names follow the scopes the reference
opens (``gnn_layer_%i`` models/sparse_graph_model.py:177; ``Edge_%i_Weight`` gnns/rgcn.py:74, ggnn.py:64, rgat.py:73,
gnn_film.py:73; ``Edge_%i_Attention_Parameters`` rgat.py:76; ``Edge_%i_FiLM_Computations`` gnn_film.py:78;
``Edge_%i_MLP`` gnn_edge_mlp.py:77, rgin.py:96; ``Edge_%i_Channel_%i_Weight_Computation`` rgdcn.py:104; ``Aggregation_MLP`` rgin.py:81; ``Dense`` sparse_graph_model.py:199;
Keras / tf.layers auto-names ``dense``, ``dense_1``, ``LayerNorm``, ``gru_cell``, ``simple_rnn_cell``).  TensorFlow is
not available in this build environment; the names were checked by running the reference's own scaffold, heads and
``save_model`` under tests/tf1_shim (tests/test_reference_model_pin.py: scopes and explicit layer names are the reference's,
the auto-numbering is the shim's restatement of TF 1.13's rules).  Matching is done on the ``gnn_layer_<i>`` component and
the components after it, whatever precedes them, and every variable that was not recognised is reported instead of being
dropped silently.
"""
import io
import pickle
import re
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

_OPTIMIZER_SLOTS = re.compile(r"(^|/)(Adam(_\d+)?|RMSProp(_\d+)?|Momentum|beta[12]_power)(:\d+)?$")
_LAYER = re.compile(r"(^|/)gnn_layer_(\d+)/(.*)$")
_AUTO = re.compile(r"^(dense|LayerNorm)(?:_(\d+))?$")


class _Placeholder:
    """Stands in for classes of the reference (or its dependencies) that are not importable here -- e.g. the
    dpu_utils Vocabulary inside VarMisuse task metadata.  Only the weights and the plain-dict params are used."""

    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k

    def __setstate__(self, state):
        self.state = state


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except Exception:
            return type(name, (_Placeholder,), {"__module__": module})


class ReferenceCheckpoint:
    def __init__(self, data: Dict[str, Any]):
        self.model_class: str = data.get("model_class", "")
        self.task_class: str = data.get("task_class", "")
        self.model_params: Dict[str, Any] = dict(data.get("model_params", {}))
        self.task_params: Dict[str, Any] = dict(data.get("task_params", {}))
        self.task_metadata = data.get("task_metadata")
        self.weights: Dict[str, np.ndarray] = {k: np.asarray(v) for k, v in data.get("weights", {}).items()}


def load_reference_checkpoint(path_or_bytes) -> ReferenceCheckpoint:
    """Read a ``*_best_model.pickle`` written by Sparse_Graph_Model.save_model."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        data = _TolerantUnpickler(io.BytesIO(path_or_bytes)).load()
    else:
        with open(path_or_bytes, "rb") as f:
            data = _TolerantUnpickler(f).load()
    if not isinstance(data, dict) or "weights" not in data:
        raise ValueError("not a reference model snapshot: expected a dict with a 'weights' entry")
    return ReferenceCheckpoint(data)


def save_reference_checkpoint(path: str, model_class: str, task_class: str, model_params: Dict, task_params: Dict,
                              weights: Dict[str, np.ndarray], task_metadata: Any = None) -> None:
    """The structure of Sparse_Graph_Model.save_model (models/sparse_graph_model.py:98-108)."""
    data = {"model_class": model_class, "task_class": task_class, "model_params": dict(model_params),
            "task_params": dict(task_params), "task_metadata": task_metadata if task_metadata is not None else {},
            "weights": {k: np.asarray(v) for k, v in weights.items()}}
    with open(path, "wb") as f:
        pickle.dump(data, f, pickle.HIGHEST_PROTOCOL)


def _strip(name: str) -> str:
    return name[:-2] if name.endswith(":0") else name


def _auto_index(component: str) -> Optional[Tuple[str, int]]:
    m = _AUTO.match(component)
    return (m.group(1), int(m.group(2) or 0)) if m else None


def sort_variables(weights: Dict[str, np.ndarray]) -> Dict[str, Any]:
    """Sort tf variable names into {"layers": [per-layer weights dict], "outside": {...}, "unused": [...]}.

    Per layer the dictionary uses the keys of this package's layer functions:
      edge_weights / attention / film_weights           lists indexed by edge type
      edge_mlps                                         per edge type, kernels in creation order (dense, dense_1, ...)
      channel_weights                                   per edge type, kernels indexed by channel (RGDCN)
      aggr_mlp                                          kernels in creation order
      ln_gamma / ln_beta                                lists indexed by timestep (LayerNorm, LayerNorm_1, ...)
      cell {"kind", "kernel", "recurrent_kernel", "bias"}
      inter_dense, inter_ln_gamma, inter_ln_beta        the scaffold's per-layer extras (sparse_graph_model.py:192-200)
    Which LayerNorm is the scaffold's: the layer functions create one per timestep FIRST, so with T timesteps the
    (T+1)-th LayerNorm of a gnn_layer scope is the inter-layer norm; callers pass ``num_timesteps`` to
    ``split_layer_norms`` to separate them.
    """
    layers: Dict[int, Dict[str, Any]] = {}
    outside: Dict[str, np.ndarray] = {}
    unused: List[str] = []

    def put_indexed(d, key, idx, value):
        d.setdefault(key, {})[idx] = value

    for full_name, value in weights.items():
        name = _strip(full_name)
        if _OPTIMIZER_SLOTS.search(name):
            continue
        m = _LAYER.search(name)
        if not m:
            outside[name] = value
            continue
        layer = layers.setdefault(int(m.group(2)), {})
        parts = m.group(3).split("/")
        head = parts[0]
        cm = re.match(r"^Edge_(\d+)_Channel_(\d+)_Weight_Computation$", head)        # gnns/rgdcn.py:104
        if cm and parts[1:] == ["kernel"]:
            layer.setdefault("channel_weights", {}).setdefault(int(cm.group(1)), {})[int(cm.group(2))] = value
            continue
        em = re.match(r"^Edge_(\d+)_(Weight|Attention_Parameters|FiLM_Computations|MLP)$", head)
        if em:
            t, kind = int(em.group(1)), em.group(2)
            if kind == "Weight" and parts[1:] == ["kernel"]:
                put_indexed(layer, "edge_weights", t, value)
            elif kind == "Attention_Parameters" and len(parts) == 1:
                put_indexed(layer, "attention", t, value)
            elif kind == "FiLM_Computations" and parts[1:] == ["kernel"]:
                put_indexed(layer, "film_weights", t, value)
            elif kind == "MLP" and len(parts) == 3 and parts[2] == "kernel" and _auto_index(parts[1]):
                layer.setdefault("edge_mlps", {}).setdefault(t, {})[_auto_index(parts[1])[1]] = value
            else:
                unused.append(full_name)
        elif head == "Aggregation_MLP" and len(parts) == 3 and parts[2] == "kernel" and _auto_index(parts[1]):
            put_indexed(layer, "aggr_mlp", _auto_index(parts[1])[1], value)
        elif _auto_index(head) and _auto_index(head)[0] == "LayerNorm" and parts[1:] in (["gamma"], ["beta"]):
            put_indexed(layer, "ln_" + parts[1], _auto_index(head)[1], value)
        elif head in ("gru_cell", "simple_rnn_cell") and len(parts) == 2 and parts[1] in ("kernel", "recurrent_kernel", "bias"):
            cell = layer.setdefault("cell", {"kind": "gru" if head == "gru_cell" else "rnn"})
            cell[parts[1]] = value
        elif head == "Dense" and parts[1:] == ["kernel"]:
            layer["inter_dense"] = value
        else:
            unused.append(full_name)

    def as_list(d: Dict[int, Any]) -> List[Any]:
        return [d[i] for i in sorted(d)]

    out_layers = []
    for i in sorted(layers):
        layer = layers[i]
        for key in ("edge_weights", "attention", "film_weights", "aggr_mlp", "ln_gamma", "ln_beta"):
            if key in layer:
                layer[key] = as_list(layer[key])
        for key in ("edge_mlps", "channel_weights"):
            if key in layer:
                layer[key] = [as_list(layer[key][t]) for t in sorted(layer[key])]
        out_layers.append(layer)
    return {"layers": out_layers, "layer_indices": sorted(layers), "outside": outside, "unused": unused}


def split_layer_norms(layer: Dict[str, Any], num_timesteps: int) -> Dict[str, Any]:
    """Separate the layer function's per-timestep LayerNorms from the scaffold's inter-layer LayerNorm."""
    layer = dict(layer)
    for key in ("ln_gamma", "ln_beta"):
        vals = layer.get(key)
        if vals is not None and len(vals) > num_timesteps:
            layer["inter_" + key] = vals[num_timesteps]
            layer[key] = vals[:num_timesteps]
    return layer


_QM9_HEAD = re.compile(r"(^|/)out_layer_task(\d+)/(regression_gate|regression)/dense/(kernel|bias)$")


def scaffold_variables(outside: Dict[str, np.ndarray], feature_size: int, hidden_size: int) -> Dict[str, np.ndarray]:
    """The variables created outside the gnn_layer scopes: the bias-free input projection
    (models/sparse_graph_model.py:166-170, absent when the feature size equals hidden_size) and the task head.
    * PPI head (tasks/ppi_task.py:176-179): one Keras Dense with bias.  Unnamed Keras layers are numbered per graph in
      creation order (dense, dense_1, ...): the projection, if any, is ``graph_model/dense``, the head the next one.
    * QM9 head (tasks/qm9_task.py:162-176): per task id ``out_layer_task<id>/regression_gate/dense/{kernel,bias}`` (gate on
      [h | x0]) and ``out_layer_task<id>/regression/dense/{kernel,bias}`` -> "qm9_heads": {task id: {"gate_kernel",
      "gate_bias", "kernel", "bias"}}.
    Names verified against the reference's own scaffold run under tests/tf1_shim (tests/test_reference_model_pin.py)."""
    dense: Dict[int, Dict[str, np.ndarray]] = {}
    qm9: Dict[int, Dict[str, np.ndarray]] = {}
    rest = {}
    for name, value in outside.items():
        parts = _strip(name).split("/")
        qm = _QM9_HEAD.search(_strip(name))
        if qm:
            key = ("gate_" if qm.group(3) == "regression_gate" else "") + qm.group(4)
            qm9.setdefault(int(qm.group(2)), {})[key] = value
            continue
        ai = _auto_index(parts[-2]) if len(parts) >= 2 else None
        if ai and ai[0] == "dense" and parts[-1] in ("kernel", "bias"):
            dense.setdefault(ai[1], {})[parts[-1]] = value
        else:
            rest[name] = value
    out: Dict[str, Any] = {"other": rest, "qm9_heads": qm9}
    order = sorted(dense)
    if feature_size != hidden_size and order:
        first = dense[order[0]]
        if "bias" not in first and first["kernel"].shape == (feature_size, hidden_size):
            out["projection"] = first["kernel"]
            order = order[1:]
    out["head"] = [dense[i] for i in order]
    return out


def _auto_name(base: str, i: int) -> str:
    return base if i == 0 else "%s_%d" % (base, i)


def layer_to_variables(layer: Dict[str, Any], prefix: str) -> Dict[str, Any]:
    """The opposite of sort_variables for ONE gnn_layer scope: a layer function's ``weights=`` dictionary (plus the
    scaffold's inter_ln_gamma / inter_ln_beta / inter_dense extras) -> {tf variable name: value}.  The scaffold's LayerNorm
    is created after the layer function's own per-timestep ones, so it is LayerNorm_<number of own norms> (A.11)."""
    out: Dict[str, Any] = {}
    for t, k in enumerate(layer.get("edge_weights") or []):
        out[prefix + "Edge_%d_Weight/kernel:0" % t] = k
    for t, a in enumerate(layer.get("attention") or []):
        out[prefix + "Edge_%d_Attention_Parameters:0" % t] = a
    for t, k in enumerate(layer.get("film_weights") or []):
        out[prefix + "Edge_%d_FiLM_Computations/kernel:0" % t] = k
    for t, kernels in enumerate(layer.get("edge_mlps") or []):
        for j, k in enumerate(kernels):
            out[prefix + "Edge_%d_MLP/%s/kernel:0" % (t, _auto_name("dense", j))] = k
    for j, k in enumerate(layer.get("aggr_mlp") or []):
        out[prefix + "Aggregation_MLP/%s/kernel:0" % _auto_name("dense", j)] = k
    for t, per_channel in enumerate(layer.get("channel_weights") or []):
        for c, k in enumerate(per_channel):
            out[prefix + "Edge_%d_Channel_%d_Weight_Computation/kernel:0" % (t, c)] = k
    gammas, betas = list(layer.get("ln_gamma") or []), list(layer.get("ln_beta") or [])
    if layer.get("inter_ln_gamma") is not None:
        gammas.append(layer["inter_ln_gamma"])
        betas.append(layer["inter_ln_beta"])
    for i, (g, b) in enumerate(zip(gammas, betas)):
        out[prefix + "%s/gamma:0" % _auto_name("LayerNorm", i)] = g
        out[prefix + "%s/beta:0" % _auto_name("LayerNorm", i)] = b
    cell = layer.get("cell")
    if cell is not None:
        scope = "gru_cell" if str(cell.get("kind", "gru")).lower() == "gru" else "simple_rnn_cell"
        for key in ("kernel", "recurrent_kernel", "bias"):
            out[prefix + "%s/%s:0" % (scope, key)] = cell[key]
    if layer.get("inter_dense") is not None:
        out[prefix + "Dense/kernel:0"] = layer["inter_dense"]
    return out


def model_to_variables(projection, layers: List[Dict[str, Any]], task: str, head, task_ids=()) -> Dict[str, Any]:
    """Whole model -> {tf variable name: value} as Sparse_Graph_Model.save_model would list them: ``graph_model/dense`` (the
    projection, when there is one), ``graph_model/gnn_layer_<i>/...``, then the head -- PPI: the next unnamed Keras Dense
    (``dense_1``, or ``dense`` without a projection; tasks/ppi_task.py:176-179); QM9: ``out_layer_task<id>/regression_gate/dense``
    and ``.../regression/dense`` (tasks/qm9_task.py:162-176)."""
    out: Dict[str, Any] = {}
    if projection is not None:
        out["graph_model/dense/kernel:0"] = projection
    for i, layer in enumerate(layers):
        out.update(layer_to_variables(layer, "graph_model/gnn_layer_%d/" % i))
    if task == "ppi":
        name = "dense_1" if projection is not None else "dense"
        out[name + "/kernel:0"], out[name + "/bias:0"] = head["kernel"], head["bias"]
    else:
        for hd, t in zip(head, task_ids):
            p = "out_layer_task%d/" % t
            out[p + "regression_gate/dense/kernel:0"], out[p + "regression_gate/dense/bias:0"] = hd["gate_kernel"], hd["gate_bias"]
            out[p + "regression/dense/kernel:0"], out[p + "regression/dense/bias:0"] = hd["kernel"], hd["bias"]
    return out


def rgcn_ppi_reference_names(num_layers: int, num_edge_types: int, has_projection: bool, inter_dense_layers,
                             prefix: str = "") -> Dict[str, str]:
    """Parameter name of RGCNPPIModel -> tf variable name the reference would use (A.11), for writing snapshots."""
    names = {}
    k = 0
    if has_projection:
        names["projection"] = prefix + "dense/kernel:0"
        k = 1
    for l in range(num_layers):
        for t in range(num_edge_types):
            names["edge_weights.%d" % (l * num_edge_types + t)] = prefix + "gnn_layer_%d/Edge_%d_Weight/kernel:0" % (l, t)
        if l in inter_dense_layers:
            names["inter_dense.%d" % l] = prefix + "gnn_layer_%d/Dense/kernel:0" % l
    head = "dense_%d" % k if k else "dense"
    names["out_kernel"] = prefix + head + "/kernel:0"
    names["out_bias"] = prefix + head + "/bias:0"
    return names
