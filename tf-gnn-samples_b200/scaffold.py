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
            return torch.optim.RMSprop(self.parameters(), lr=p["learning_rate"], alpha=p["learning_rate_decay"],
                                       momentum=p["momentum"], eps=1e-10)
        if name == "adam":
            return torch.optim.Adam(self.parameters(), lr=p["learning_rate"], eps=1e-8)
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

    def train_step_async(self, optimizer, features, plan, num_incoming, labels) -> Dict[str, torch.Tensor]:
        """One step of __make_train_step: gradients of the per-node loss, per-tensor clip_by_norm, apply.  Nothing in
        here waits for the device; the metrics come back as device tensors."""
        self.train()
        optimizer.zero_grad(set_to_none=True)
        m = self.task_metrics(self(features, plan, num_incoming), labels)
        m["loss"].backward()
        self.clip_gradients_()
        optimizer.step()
        return {k: v.detach() for k, v in m.items()}

    def train_step(self, optimizer, features, plan, num_incoming, labels) -> Dict[str, float]:
        """train_step_async + one device->host read of the three metrics."""
        m = self.train_step_async(optimizer, features, plan, num_incoming, labels)
        keys = list(m)
        vals = torch.stack([m[k].float() for k in keys]).tolist()
        return dict(zip(keys, vals))
