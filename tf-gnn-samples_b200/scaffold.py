"""Torch scaffold around the engine's RGCN layer: the forward / train step of RGCN_Model on the PPI task
(SURVEY.md 8f rank 2 -- a thin host-side mirror so the op can be exercised the way the reference's
train.py / test.py use it; the tasks, data loading, CLI and logging of the reference stay out of scope).

Mirrors:
  * Sparse_Graph_Model.__build_graph_propagation_model (models/sparse_graph_model.py:162-202): bias-free input
    projection with graph_model_activation_function (skipped when the feature size equals hidden_size), per layer
    input dropout, residual every k layers, the GNN layer, optional inter-layer LayerNorm, and the extra
    ``activation(Dense)`` after every ``graph_dense_between_every_num_gnn_layers``-th layer (layer 0 included);
  * RGCN_Model.default_params / _apply_gnn_layer (models/rgcn_model.py:12-44);
  * PPI_Task.make_task_output_model (tasks/ppi_task.py:165-195): Dense(121, bias) logits, sigmoid cross entropy
    summed over labels and nodes, loss = total / num_nodes, micro-F1 (utils/utils.py:61-74);
  * Sparse_Graph_Model.__make_train_step (models/sparse_graph_model.py:227-260): Adam / RMSProp / SGD with
    per-tensor tf.clip_by_norm(clamp_gradient_norm).
The GNN layers run on the engine (sparse_rgcn_layer, with its own backward); the scaffold's own small Dense layers
are plain torch matmuls (library GEMMs outside the hot path).
"""
from typing import Dict, List, Optional

import numpy as np
import torch

from .engine import GraphPlan
from .gnns import sparse_rgcn_layer
from .ops import dense as engine_dense
from .weights import glorot_uniform
from .tf_optimizers import TF1Adam, TF1RMSProp

_ACT = {"tanh": torch.tanh, "relu": torch.relu, "linear": lambda x: x, None: lambda x: x,
        "elu": torch.nn.functional.elu, "gelu": torch.nn.functional.gelu,
        "leaky_relu": lambda x: torch.nn.functional.leaky_relu(x, 0.2), "selu": torch.selu}


def _matmul(x: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    """x @ kernel on the engine's tensor-core GEMM (forward and both gradients).  Widths that are not multiples
    of 4 (PPI: 50 input features, 121 labels) are zero-padded to the next multiple -- exact, the padding
    contributes zeros -- so that rows stay 16-byte aligned for the kernels."""
    k, n = kernel.shape
    pk, pn = (-k) % 4, (-n) % 4
    if pk:
        x = torch.nn.functional.pad(x, (0, pk))
    if pk or pn:
        kernel = torch.nn.functional.pad(kernel, (0, pn, 0, pk))
    y = engine_dense(x, kernel)
    return y[:, :n] if pn else y


def all_reduce_gradients_(parameters, group=None) -> None:
    """Data-parallel training over graph-boundary shards: SUM the gradients of every parameter over the
    ranks with ONE collective on a flattened buffer (NCCL on GPUs, gloo in the CPU tests).  Callers scale their local loss
    by the GLOBAL normaliser (all nodes / all graphs of the union batch), so the sum is exactly the single-device gradient
    of the union batch.  A parameter without a local gradient (e.g. the kernel of an edge type absent from this shard)
    contributes zeros."""
    import torch.distributed as dist
    params = [q for q in parameters if q.requires_grad]
    if not params or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for q in params:
        if q.grad is None:
            q.grad = torch.zeros_like(q)
    flat = torch.cat([q.grad.reshape(-1) for q in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for q in params:
        n = q.numel()
        q.grad.copy_(flat[off:off + n].view_as(q.grad))
        off += n


def global_count(local_count: int, device, group=None) -> float:
    """Sum of a per-rank count (nodes or graphs of the shard) over the ranks."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(local_count)
    t = torch.tensor([float(local_count)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def rgcn_ppi_default_params() -> Dict:
    """Sparse_Graph_Model.default_params overlaid with RGCN_Model.default_params and the README's PPI run
    (hidden 256, 3 layers: README.md:29-35)."""
    return {
        "hidden_size": 256, "graph_num_layers": 3, "graph_num_timesteps_per_layer": 1,
        "graph_layer_input_dropout_keep_prob": 1.0, "graph_dense_between_every_num_gnn_layers": 10000,
        "graph_model_activation_function": "tanh", "graph_residual_connection_every_num_layers": 10000,
        "graph_inter_layer_norm": False, "graph_activation_function": "ReLU",
        "message_aggregation_function": "sum", "optimizer": "Adam", "learning_rate": 0.001,
        "learning_rate_decay": 0.98, "momentum": 0.85, "clamp_gradient_norm": 1.0, "random_seed": 0,
    }


class RGCNPPIModel(torch.nn.Module):
    """RGCN_Model on the PPI task.  Parameters use the reference's variable names."""

    def __init__(self, num_edge_types: int = 3, feature_size: int = 50, num_labels: int = 121,
                 params: Optional[Dict] = None, device="cuda"):
        super().__init__()
        self.params = dict(rgcn_ppi_default_params(), **(params or {}))
        H = self.params["hidden_size"]
        rng = np.random.default_rng(self.params["random_seed"])
        P = lambda a: torch.nn.Parameter(torch.as_tensor(a, dtype=torch.float32, device=device))
        self.feature_size, self.num_edge_types = feature_size, num_edge_types
        self.projection = P(glorot_uniform(rng, feature_size, H)) if feature_size != H else None   # :165-170
        self.edge_weights = torch.nn.ParameterList()
        self.inter_dense = torch.nn.ParameterDict()
        self.inter_ln = torch.nn.ParameterDict()
        for l in range(self.params["graph_num_layers"]):
            for t in range(num_edge_types):
                self.edge_weights.append(P(glorot_uniform(rng, H, H)))                              # gnn_layer_l/Edge_t_Weight/kernel
            if self.params["graph_inter_layer_norm"]:
                self.inter_ln["g%d" % l], self.inter_ln["b%d" % l] = P(np.ones(H)), P(np.zeros(H))
            if l % self.params["graph_dense_between_every_num_gnn_layers"] == 0:                    # :194 (fires for layer 0)
                self.inter_dense[str(l)] = P(glorot_uniform(rng, H, H))                             # gnn_layer_l/Dense/kernel
        self.out_kernel = P(glorot_uniform(rng, H, num_labels))                                    # ppi_task.py:176-179
        self.out_bias = P(np.zeros(num_labels))

    def layer_weights(self, l: int) -> Dict[str, List[torch.Tensor]]:
        L = self.num_edge_types
        return {"edge_weights": [self.edge_weights[l * L + t] for t in range(L)]}

    # ---- reference snapshots (models/sparse_graph_model.py:91-126) ----
    def reference_variable_names(self) -> Dict[str, str]:
        from .checkpoint import rgcn_ppi_reference_names
        names = rgcn_ppi_reference_names(self.params["graph_num_layers"], self.num_edge_types, self.projection is not None,
                                         [int(k) for k in self.inter_dense])
        for k in self.inter_ln:                                     # "g<l>" / "b<l>"
            names["inter_ln." + k] = "gnn_layer_%d/LayerNorm/%s:0" % (int(k[1:]), "gamma" if k[0] == "g" else "beta")
        return names

    def to_reference_weights(self) -> Dict[str, np.ndarray]:
        """{tf variable name: array} as Sparse_Graph_Model.save_model stores it."""
        names = self.reference_variable_names()
        return {names[n]: p.detach().cpu().numpy() for n, p in self.named_parameters()}

    def load_reference_weights(self, weights: Dict[str, np.ndarray]) -> List[str]:
        """Assign a reference snapshot's ``weights`` dictionary by variable name (load_weights, :110-126).
        Returns the names that were not used; raises if a parameter of this model has no saved value or a
        different shape (the reference would re-initialise it silently -- here that is an error)."""
        from .checkpoint import scaffold_variables, sort_variables, split_layer_norms
        srt = sort_variables(weights)
        H, L = self.params["hidden_size"], self.num_edge_types
        outside = scaffold_variables(srt["outside"], self.feature_size, H)
        found: Dict[str, np.ndarray] = {}
        if self.projection is not None and "projection" in outside:
            found["projection"] = outside["projection"]
        for idx, layer in zip(srt["layer_indices"], srt["layers"]):
            layer = split_layer_norms(layer, 0)                     # RGCN itself has no LayerNorm
            for t, w in enumerate(layer.get("edge_weights", [])):
                found["edge_weights.%d" % (idx * L + t)] = w
            if "inter_dense" in layer:
                found["inter_dense.%d" % idx] = layer["inter_dense"]
            if "inter_ln_gamma" in layer:
                found["inter_ln.g%d" % idx], found["inter_ln.b%d" % idx] = layer["inter_ln_gamma"], layer["inter_ln_beta"]
        heads = [h for h in outside["head"] if "bias" in h]
        if heads:
            found["out_kernel"], found["out_bias"] = heads[-1]["kernel"], heads[-1]["bias"]
        used = set()
        with torch.no_grad():
            for n, p in self.named_parameters():
                if n not in found:
                    raise KeyError("reference snapshot has no value for %s" % n)
                v = np.asarray(found[n], dtype=np.float32)
                if tuple(v.shape) != tuple(p.shape):
                    raise ValueError("%s: snapshot shape %s != model shape %s" % (n, v.shape, tuple(p.shape)))
                p.copy_(torch.as_tensor(v))
                used.add(n)
        return list(srt["unused"]) + sorted(outside["other"])

    def num_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters())

    def node_representations(self, features: torch.Tensor, plan: GraphPlan, num_incoming: torch.Tensor) -> torch.Tensor:
        p = self.params
        act = _ACT[p["graph_model_activation_function"].lower() if p["graph_model_activation_function"] else None]
        cur = features if self.projection is None else act(_matmul(features, self.projection))
        last_residual = torch.zeros_like(cur)
        keep = p["graph_layer_input_dropout_keep_prob"]
        for l in range(p["graph_num_layers"]):
            if self.training and keep < 1.0:
                cur = torch.nn.functional.dropout(cur, p=1.0 - keep)                                # :179
            if l % p["graph_residual_connection_every_num_layers"] == 0:                            # :180-185
                t = cur
                if l > 0:
                    cur = (cur + last_residual) / 2
                last_residual = t
            cur = sparse_rgcn_layer(cur, plan, num_incoming, p["hidden_size"],                      # :187-191 -> rgcn_model.py:36-44
                                    num_timesteps=p["graph_num_timesteps_per_layer"],
                                    activation_function=p["graph_activation_function"],
                                    message_aggregation_function=p["message_aggregation_function"],
                                    weights=self.layer_weights(l))
            if p["graph_inter_layer_norm"]:                                                         # :192-193 (eps 1e-12)
                cur = torch.nn.functional.layer_norm(cur, (p["hidden_size"],), self.inter_ln["g%d" % l], self.inter_ln["b%d" % l], 1e-12)
            if str(l) in self.inter_dense:                                                          # :194-200
                cur = act(_matmul(cur, self.inter_dense[str(l)]))
        return cur

    def forward(self, features, plan, num_incoming):
        return _matmul(self.node_representations(features, plan, num_incoming), self.out_kernel) + self.out_bias

    def task_metrics(self, logits: torch.Tensor, labels: torch.Tensor) -> Dict[str, torch.Tensor]:
        """tasks/ppi_task.py:181-195.  round(sigmoid(x)) == 1 exactly when x > 0 (round-half-even sends 0.5 to 0),
        so the predictions are taken from the sign of the logits and the three counts come from one reduction."""
        total = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction="sum")
        pred, lab = logits > 0, labels > 0.5
        counts = torch.stack([pred & lab, pred & ~lab, ~pred & lab]).sum(dim=(1, 2)).double()   # tp, fp, fn
        tp, fp, fn = counts[0], counts[1], counts[2]
        precision, recall = tp / (tp + fp), tp / (tp + fn)
        return {"loss": total / logits.shape[0], "total_loss": total,
                "f1_score": (2 * precision * recall / (precision + recall)).float()}

    def make_optimizer(self) -> torch.optim.Optimizer:
        p, name = self.params, self.params["optimizer"].lower()
        if name == "sgd":
            return torch.optim.SGD(self.parameters(), lr=p["learning_rate"])
        if name == "rmsprop":
            return TF1RMSProp(self.parameters(), lr=p["learning_rate"], decay=p["learning_rate_decay"],
                              momentum=p["momentum"], epsilon=1e-10)         # the TF 1.13 rule, not torch.optim.RMSprop
        if name == "adam":
            return TF1Adam(self.parameters(), lr=p["learning_rate"], epsilon=1e-8)
        raise Exception('Unknown optimizer "%s".' % p["optimizer"])

    def clip_gradients_(self) -> None:
        """tf.clip_by_norm per tensor (:255-258): g * clip / max(||g||, clip) -- evaluated on the device, no host sync."""
        grads = [q.grad for q in self.parameters() if q.grad is not None]
        if not grads:
            return
        clip = float(self.params["clamp_gradient_norm"])
        scales = torch._foreach_norm(grads)
        torch._foreach_clamp_min_(scales, clip)
        torch._foreach_reciprocal_(scales)
        torch._foreach_mul_(scales, clip)
        torch._foreach_mul_(grads, scales)

    def train_step_async(self, optimizer, features, plan, num_incoming, labels, group=None,
                         global_num_nodes: Optional[float] = None) -> Dict[str, torch.Tensor]:
        """One step of __make_train_step: gradients of the per-node loss, per-tensor clip_by_norm, apply.  Nothing in
        here waits for the device; the metrics come back as device tensors.  With ``global_num_nodes`` (the node count of
        the union batch over all ranks) the step is data-parallel: the local loss is total / global_num_nodes and the
        gradients are summed over ``group`` before clipping, which reproduces the single-device step on the union batch."""
        self.train()
        optimizer.zero_grad(set_to_none=True)
        m = self.task_metrics(self(features, plan, num_incoming), labels)
        if global_num_nodes is None:
            m["loss"].backward()
        else:
            (m["total_loss"] / float(global_num_nodes)).backward()
            all_reduce_gradients_(self.parameters(), group)
        self.clip_gradients_()
        optimizer.step()
        return {k: v.detach() for k, v in m.items()}

    def train_step(self, optimizer, features, plan, num_incoming, labels) -> Dict[str, float]:
        """train_step_async + one device->host read of the three metrics."""
        m = self.train_step_async(optimizer, features, plan, num_incoming, labels)
        keys = list(m)
        vals = torch.stack([m[k].float() for k in keys]).tolist()
        return dict(zip(keys, vals))


# =====================================================================================================================
# General scaffold: any of the six layer families under either task head
# =====================================================================================================================
_BASE_DEFAULTS = {                                             # Sparse_Graph_Model.default_params (models/sparse_graph_model.py:33-55)
    "max_nodes_in_batch": 50000, "graph_num_layers": 8, "graph_num_timesteps_per_layer": 1,
    "graph_layer_input_dropout_keep_prob": 0.8, "graph_dense_between_every_num_gnn_layers": 1,
    "graph_model_activation_function": "tanh", "graph_residual_connection_every_num_layers": 2,
    "graph_inter_layer_norm": False, "optimizer": "Adam", "learning_rate": 0.001, "learning_rate_decay": 0.98,
    "momentum": 0.85, "clamp_gradient_norm": 1.0, "random_seed": 0, "lr_for_num_graphs_per_batch": None,
}
_MODEL_DEFAULTS = {                                            # <X>_Model.default_params overlays (models/*_model.py)
    "rgcn": {"hidden_size": 128, "graph_activation_function": "ReLU", "message_aggregation_function": "sum",
             "graph_layer_input_dropout_keep_prob": 1.0, "graph_dense_between_every_num_gnn_layers": 10000,
             "graph_residual_connection_every_num_layers": 10000},
    "ggnn": {"hidden_size": 128, "graph_rnn_cell": "GRU", "graph_activation_function": "tanh",
             "message_aggregation_function": "sum", "graph_layer_input_dropout_keep_prob": 1.0,
             "graph_dense_between_every_num_gnn_layers": 10000, "graph_residual_connection_every_num_layers": 10000},
    "rgat": {"hidden_size": 128, "num_heads": 4, "graph_activation_function": "tanh",
             "graph_layer_input_dropout_keep_prob": 1.0, "graph_dense_between_every_num_gnn_layers": 10000,
             "graph_residual_connection_every_num_layers": 10000},
    "rgin": {"hidden_size": 128, "graph_activation_function": "ReLU", "message_aggregation_function": "sum",
             "graph_dense_between_every_num_gnn_layers": 10000, "graph_inter_layer_norm": True,
             "use_target_state_as_input": False, "graph_num_edge_MLP_hidden_layers": 1,
             "graph_num_aggr_MLP_hidden_layers": None},
    "gnn-edge-mlp": {"max_nodes_in_batch": 25000, "hidden_size": 128, "graph_activation_function": "gelu",
                     "message_aggregation_function": "sum", "graph_inter_layer_norm": True,
                     "use_target_state_as_input": True, "num_edge_hidden_layers": 1},
    "gnn-film": {"hidden_size": 128, "graph_activation_function": "ReLU", "message_aggregation_function": "sum",
                 "normalize_messages_by_num_incoming": False},
    "rgdcn": {"max_nodes_in_batch": 25000, "hidden_size": 128, "num_channels": 8, "use_full_state_for_channel_weights": False,
              "tie_channel_weights": False, "graph_activation_function": "ReLU", "message_aggregation_function": "sum",
              "graph_inter_layer_norm": True},
}
_MODEL_NAMES = {                                               # name_to_model_class (utils/model_utils.py:31-55): name -> (family, extra params)
    "ggnn": ("ggnn", {}), "ggnn_model": ("ggnn", {}),
    "gnn_edge_mlp": ("gnn-edge-mlp", {}), "gnn-edge-mlp": ("gnn-edge-mlp", {}), "gnn_edge_mlp_model": ("gnn-edge-mlp", {}),
    "gnn_edge_mlp0": ("gnn-edge-mlp", {"num_edge_hidden_layers": 0}), "gnn-edge-mlp0": ("gnn-edge-mlp", {"num_edge_hidden_layers": 0}),
    "gnn_edge_mlp0_model": ("gnn-edge-mlp", {"num_edge_hidden_layers": 0}),
    "gnn_edge_mlp1": ("gnn-edge-mlp", {"num_edge_hidden_layers": 1}), "gnn-edge-mlp1": ("gnn-edge-mlp", {"num_edge_hidden_layers": 1}),
    "gnn_edge_mlp1_model": ("gnn-edge-mlp", {"num_edge_hidden_layers": 1}),
    "gnn_film": ("gnn-film", {}), "gnn-film": ("gnn-film", {}), "gnn_film_model": ("gnn-film", {}),
    "rgat": ("rgat", {}), "rgat_model": ("rgat", {}), "rgcn": ("rgcn", {}), "rgcn_model": ("rgcn", {}),
    "rgdcn": ("rgdcn", {}), "rgdcn_model": ("rgdcn", {}), "rgin": ("rgin", {}), "rgin_model": ("rgin", {}),
}


def resolve_model_name(model: str):
    """(layer family, extra params) for a model name, case-insensitively, exactly the names name_to_model_class accepts
    (``gnn_edge_mlp0`` / ``gnn-edge-mlp1`` ... fix num_edge_hidden_layers); ValueError("Unknown model type '<lowered name>'")."""
    name = model.lower()
    if name not in _MODEL_NAMES:
        raise ValueError("Unknown model type '%s'" % name)
    return _MODEL_NAMES[name]


_LAYERS_WITH_OWN_LN = ("rgin", "gnn-edge-mlp", "gnn-film")      # one LayerNorm per timestep inside the layer function


def model_default_params(model: str) -> Dict:
    """name_to_model_class(...)[0].default_params() of the reference (utils/model_utils.py:30-55)."""
    kind, extra = resolve_model_name(model)
    return {**_BASE_DEFAULTS, **_MODEL_DEFAULTS[kind], **extra}


class SparseGraphModel(torch.nn.Module):
    """Sparse_Graph_Model with one of the six GNN layer families and a PPI (node classification) or QM9 (graph
    regression) head.  Inference runs the fused layer kernels; under autograd the layers take their differentiable
    paths (gnns/_train.py), so ``train_step`` works for every family."""

    def __init__(self, model: str, task: str, num_edge_types: int, feature_size: int, params: Optional[Dict] = None,
                 num_labels: int = 121, task_ids=(0,), device="cuda"):
        super().__init__()
        from . import weights as W
        self.kind, _ = resolve_model_name(model)
        self.task = task.lower()
        if self.task not in ("ppi", "qm9"):
            raise ValueError("Unknown task type '%s'" % task)
        self.params = dict(model_default_params(model), **(params or {}))
        p = self.params
        H, L, T = p["hidden_size"], num_edge_types, p["graph_num_timesteps_per_layer"]
        self.num_edge_types, self.feature_size, self.task_ids = L, feature_size, tuple(task_ids)
        self._device = torch.device(device)
        rng = np.random.default_rng(p["random_seed"])
        self._count = 0
        self.projection = self._param(glorot_uniform(rng, feature_size, H)) if feature_size != H else None
        self.layers: List[Dict] = []
        for l in range(p["graph_num_layers"]):
            seed = p["random_seed"] * 1000 + 17 * l + 1
            if self.kind == "rgcn":
                w = W.rgcn_weights(L, H, H, seed)
            elif self.kind == "ggnn":
                w = W.ggnn_weights(L, H, seed, cell=p["graph_rnn_cell"])
            elif self.kind == "rgat":
                w = W.rgat_weights(L, H, H, seed)
            elif self.kind == "gnn-film":
                w = W.film_weights(L, H, H, seed, num_timesteps=T)
            elif self.kind == "rgdcn":                          # channel_dim = hidden_size // num_channels (rgdcn_model.py:30-32)
                w = W.rgdcn_weights(L, p["num_channels"], H // p["num_channels"], p["use_full_state_for_channel_weights"],
                                    p["tie_channel_weights"], seed)
            elif self.kind == "gnn-edge-mlp":
                w = W.edge_mlp_weights(L, H, H, p["num_edge_hidden_layers"], p["use_target_state_as_input"], seed, num_timesteps=T)
            else:
                w = W.rgin_weights(L, H, H, p["graph_num_edge_MLP_hidden_layers"], p["graph_num_aggr_MLP_hidden_layers"],
                                   p["use_target_state_as_input"], seed, num_timesteps=T)
            extras = {}
            if p["graph_inter_layer_norm"]:
                extras["inter_ln_gamma"], extras["inter_ln_beta"] = np.ones(H, np.float32), np.zeros(H, np.float32)
            if l % p["graph_dense_between_every_num_gnn_layers"] == 0:
                extras["inter_dense"] = glorot_uniform(rng, H, H)
            self.layers.append(self._register("gnn_layer_%d" % l, dict(w, **extras)))
        if self.task == "ppi":
            self.head = self._register("head", {"kernel": glorot_uniform(rng, H, num_labels), "bias": np.zeros(num_labels, np.float32)})
        else:                                                    # tasks/qm9_task.py:162-176: per task a gate on [h | x0] and a transform on h
            self.head = self._register("head", [{"gate_kernel": glorot_uniform(rng, H + feature_size, 1), "gate_bias": np.zeros(1, np.float32),
                                                 "kernel": glorot_uniform(rng, H, 1), "bias": np.zeros(1, np.float32)}
                                                for _ in self.task_ids])

    # ---- parameter plumbing: nested containers of Parameters, registered under flat names ----
    def _param(self, array, name: Optional[str] = None):
        q = torch.nn.Parameter(torch.as_tensor(np.ascontiguousarray(array), dtype=torch.float32, device=self._device))
        self.register_parameter(name or "projection", q)
        return q

    def _register(self, prefix: str, obj):
        if isinstance(obj, dict):
            return {k: self._register("%s__%s" % (prefix, k), v) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return [self._register("%s__%d" % (prefix, i), v) for i, v in enumerate(obj)]
        if obj is None or isinstance(obj, str):
            return obj
        return self._param(obj, prefix)

    def num_parameters(self) -> int:
        return sum(q.numel() for q in self.parameters())

    # ---- models/<x>_model.py: _apply_gnn_layer ----
    def _apply_gnn_layer(self, cur, plan, cnt, w):
        from . import gnns as G
        p, H, T = self.params, self.params["hidden_size"], self.params["graph_num_timesteps_per_layer"]
        act = p["graph_activation_function"]
        if self.kind == "rgcn":
            return G.sparse_rgcn_layer(cur, plan, cnt, H, num_timesteps=T, activation_function=act,
                                       message_aggregation_function=p["message_aggregation_function"], weights=w)
        if self.kind == "ggnn":
            return G.sparse_ggnn_layer(cur, plan, H, num_timesteps=T, gated_unit_type=p["graph_rnn_cell"], activation_function=act,
                                       message_aggregation_function=p["message_aggregation_function"], weights=w)
        if self.kind == "rgat":
            return G.sparse_rgat_layer(cur, plan, H, num_timesteps=T, num_heads=p["num_heads"], activation_function=act, weights=w)
        if self.kind == "gnn-film":
            return G.sparse_gnn_film_layer(cur, plan, cnt, H, num_timesteps=T, activation_function=act,
                                           message_aggregation_function=p["message_aggregation_function"],
                                           normalize_by_num_incoming=p["normalize_messages_by_num_incoming"], weights=w)
        if self.kind == "rgdcn":                                  # inference only: the layer has no gradient path
            return G.sparse_rgdcn_layer(cur, plan, cnt, p["num_channels"], H // p["num_channels"], num_timesteps=T,
                                        use_full_state_for_channel_weights=p["use_full_state_for_channel_weights"],
                                        tie_channel_weights=p["tie_channel_weights"], activation_function=act,
                                        message_aggregation_function=p["message_aggregation_function"], weights=w)
        if self.kind == "gnn-edge-mlp":
            return G.sparse_gnn_edge_mlp_layer(cur, plan, cnt, H, num_timesteps=T, activation_function=act,
                                               message_aggregation_function=p["message_aggregation_function"],
                                               use_target_state_as_input=p["use_target_state_as_input"],
                                               num_edge_hidden_layers=p["num_edge_hidden_layers"], weights=w)
        return G.sparse_rgin_layer(cur, plan, H, num_timesteps=T, activation_function=act,
                                   message_aggregation_function=p["message_aggregation_function"],
                                   use_target_state_as_input=p["use_target_state_as_input"],
                                   num_edge_MLP_hidden_layers=p["graph_num_edge_MLP_hidden_layers"],
                                   num_aggr_MLP_hidden_layers=p["graph_num_aggr_MLP_hidden_layers"], weights=w)

    # ---- models/sparse_graph_model.py:162-202 ----
    def node_representations(self, features, plan: GraphPlan, num_incoming):
        p = self.params
        act = _ACT[p["graph_model_activation_function"].lower() if p["graph_model_activation_function"] else None]
        cur = features if self.projection is None else act(_matmul(features, self.projection))
        last_residual = torch.zeros_like(cur)
        keep = p["graph_layer_input_dropout_keep_prob"]
        for l, w in enumerate(self.layers):
            if self.training and keep < 1.0:
                cur = torch.nn.functional.dropout(cur, p=1.0 - keep)
            if l % p["graph_residual_connection_every_num_layers"] == 0:
                t = cur
                if l > 0:
                    cur = (cur + last_residual) / 2
                last_residual = t
            cur = self._apply_gnn_layer(cur, plan, num_incoming, w)
            if "inter_ln_gamma" in w:
                cur = torch.nn.functional.layer_norm(cur, (p["hidden_size"],), w["inter_ln_gamma"], w["inter_ln_beta"], 1e-12)
            if "inter_dense" in w:
                cur = act(_matmul(cur, w["inter_dense"]))
        return cur

    def forward(self, features, plan, num_incoming, graph_nodes_list=None, num_graphs: Optional[int] = None):
        """PPI: per-node logits [V, num_labels].  QM9: per-graph outputs [len(task_ids), G] (tasks/qm9_task.py:178-189)."""
        final = self.node_representations(features, plan, num_incoming)
        if self.task == "ppi":
            return _matmul(final, self.head["kernel"]) + self.head["bias"]
        gate_in = torch.cat([final, features], dim=-1)
        outs = []
        for hd in self.head:
            per_node = _matmul(final, hd["kernel"]) + hd["bias"]
            gated = torch.sigmoid(_matmul(gate_in, hd["gate_kernel"]) + hd["gate_bias"]) * per_node
            outs.append(torch.zeros((int(num_graphs), 1), dtype=torch.float32, device=final.device)
                        .index_add(0, graph_nodes_list.long(), gated).squeeze(-1))
        return torch.stack(outs)

    def task_metrics(self, outputs, targets) -> Dict[str, torch.Tensor]:
        if self.task == "ppi":
            return RGCNPPIModel.task_metrics(self, outputs, targets)
        err = outputs - targets                                                  # [tasks, G]
        m = {"abs_err_task%d" % t: err[i].abs().sum() for i, t in enumerate(self.task_ids)}
        m["loss"] = (0.5 * err * err).mean(dim=1).sum()                           # tasks/qm9_task.py:195-197
        m["total_loss"] = m["loss"] * outputs.shape[1]
        return m

    make_optimizer = RGCNPPIModel.make_optimizer
    clip_gradients_ = RGCNPPIModel.clip_gradients_

    def set_learning_rate_(self, optimizer, num_graphs: Optional[int]) -> None:
        """models/sparse_graph_model.py:230-238: with ``lr_for_num_graphs_per_batch`` = n the step's learning rate is
        learning_rate * num_graphs / n (evaluated in float32 there), so the rate PER GRAPH stays fixed while the packed batches
        vary in size.  Only the VarMisuse hyper-parameter files set it; None (the default) leaves the optimizer alone."""
        n = self.params.get("lr_for_num_graphs_per_batch")
        if n is None:
            return
        if num_graphs is None:
            raise ValueError("lr_for_num_graphs_per_batch is set: train_step needs num_graphs")
        lr = float(self.params["learning_rate"]) * float(np.float32(num_graphs) / np.float32(n))
        for group in optimizer.param_groups:
            group["lr"] = lr

    def train_step_async(self, optimizer, features, plan, num_incoming, targets, graph_nodes_list=None,
                         num_graphs: Optional[int] = None, group=None, global_count_: Optional[float] = None) -> Dict[str, torch.Tensor]:
        """``global_count_``: nodes (PPI) / graphs (QM9) of the union batch over all ranks -> data-parallel step
        (see RGCNPPIModel.train_step_async)."""
        self.train()
        optimizer.zero_grad(set_to_none=True)
        self.set_learning_rate_(optimizer, num_graphs)
        m = self.task_metrics(self(features, plan, num_incoming, graph_nodes_list, num_graphs), targets)
        if global_count_ is None:
            m["loss"].backward()
        else:
            (m["total_loss"] / float(global_count_)).backward()      # total_loss = loss * local count for both heads
            all_reduce_gradients_(self.parameters(), group)
        self.clip_gradients_()
        optimizer.step()
        return {k: v.detach() for k, v in m.items()}

    def train_step(self, optimizer, features, plan, num_incoming, targets, graph_nodes_list=None,
                   num_graphs: Optional[int] = None) -> Dict[str, float]:
        m = self.train_step_async(optimizer, features, plan, num_incoming, targets, graph_nodes_list, num_graphs)
        keys = list(m)
        return dict(zip(keys, torch.stack([m[k].float() for k in keys]).tolist()))

    # ---- reference snapshots ----
    def to_reference_weights(self) -> Dict[str, np.ndarray]:
        """{tf variable name: array} as Sparse_Graph_Model.save_model stores it (models/sparse_graph_model.py:91-97): what the
        reference's restore() / load_weights assigns by name.  Inverse of load_reference_weights."""
        from .checkpoint import model_to_variables
        named = model_to_variables(self.projection, self.layers, self.task, self.head, self.task_ids)
        out = {k: v.detach().cpu().numpy() for k, v in named.items()}
        out["total_num_graphs:0"] = np.zeros((), np.int64)       # the scaffold's graph counter (:143-148) is a global variable too
        return out

    def save_reference_snapshot(self, path: str, task_params: Optional[Dict] = None, task_metadata: Optional[Dict] = None) -> None:
        """Write the pickle of save_model (:98-107): model_class / task_class are the names utils/model_utils.py maps back to
        classes (RGCN, GGNN, RGAT, RGIN, RGDCN, GNN-Edge-MLP<n>, GNN-FiLM; PPI, QM9)."""
        from .checkpoint import save_reference_checkpoint
        names = {"rgcn": "RGCN", "ggnn": "GGNN", "rgat": "RGAT", "rgin": "RGIN", "rgdcn": "RGDCN", "gnn-film": "GNN-FiLM",
                 "gnn-edge-mlp": "GNN-Edge-MLP%d" % self.params.get("num_edge_hidden_layers", 1)}
        save_reference_checkpoint(path, names[self.kind], self.task.upper(), self.params, task_params or {},
                                  self.to_reference_weights(), task_metadata)

    def load_reference_weights(self, weights: Dict[str, np.ndarray]) -> List[str]:
        """Assign a reference snapshot by tf variable name (checkpoint.sort_variables); the per-layer dictionaries of
        the snapshot and of this model use the same keys.  Raises when a parameter has no saved value / another shape."""
        from .checkpoint import scaffold_variables, sort_variables, split_layer_norms
        srt = sort_variables(weights)
        T = self.params["graph_num_timesteps_per_layer"] if self.kind in _LAYERS_WITH_OWN_LN else 0
        by_index = {i: split_layer_norms(l, T) for i, l in zip(srt["layer_indices"], srt["layers"])}
        outside = scaffold_variables(srt["outside"], self.feature_size, self.params["hidden_size"])

        def assign(dst, src, path):
            if isinstance(dst, dict):
                for k, v in dst.items():
                    if k == "kind" or v is None:
                        continue
                    if not isinstance(src, dict) or k not in src:
                        raise KeyError("reference snapshot has no value for %s.%s" % (path, k))
                    assign(v, src[k], "%s.%s" % (path, k))
            elif isinstance(dst, (list, tuple)):
                if not isinstance(src, (list, tuple)) or len(src) != len(dst):
                    raise KeyError("reference snapshot: %s needs %d entries" % (path, len(dst)))
                for i, v in enumerate(dst):
                    assign(v, src[i], "%s.%d" % (path, i))
            else:
                a = np.asarray(src, dtype=np.float32)
                if tuple(a.shape) != tuple(dst.shape):
                    raise ValueError("%s: snapshot shape %s != model shape %s" % (path, a.shape, tuple(dst.shape)))
                with torch.no_grad():
                    dst.copy_(torch.as_tensor(a))

        if self.projection is not None:
            if "projection" not in outside:
                raise KeyError("reference snapshot has no input projection kernel")
            assign(self.projection, outside["projection"], "projection")
        for l, w in enumerate(self.layers):
            if l not in by_index:
                raise KeyError("reference snapshot has no variables for gnn_layer_%d" % l)
            assign(w, by_index[l], "gnn_layer_%d" % l)
        if self.task == "ppi":
            heads = [h for h in outside["head"] if "bias" in h]
            if not heads:
                raise KeyError("reference snapshot has no output layer")
            assign(self.head, heads[-1], "head")
        else:                                                    # out_layer_task<id>/{regression_gate,regression}/dense/*
            for i, t in enumerate(self.task_ids):
                if t not in outside["qm9_heads"]:
                    raise KeyError("reference snapshot has no output layer for task %d" % t)
                assign(self.head[i], outside["qm9_heads"][t], "out_layer_task%d" % t)
        return list(srt["unused"])
