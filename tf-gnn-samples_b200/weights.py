"""Weight construction with the initialisers the reference's layers use:
Keras Dense / tf.get_variable default = Glorot-uniform, GRU recurrent kernel = orthogonal, biases = 0,
layer-norm gamma = 1 / beta = 0.  Arrays are numpy float32 in Keras orientation ([in, out]); the dict
layouts are the ``weights=`` arguments of the layer functions in ``gnns/``.  Random values are only
used for parity tests and benchmarks (random-init weights of the reference architecture)."""
from typing import Dict, List, Optional

import numpy as np


def glorot_uniform(rng: np.random.Generator, fan_in: int, fan_out: int, shape=None) -> np.ndarray:
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape if shape is not None else (fan_in, fan_out)).astype(np.float32)


def orthogonal(rng: np.random.Generator, rows: int, cols: int) -> np.ndarray:
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    q = q if rows >= cols else q.T
    return np.ascontiguousarray(q[:rows, :cols]).astype(np.float32)


def rgcn_weights(L: int, d_in: int, d_out: int, seed: int = 2, use_both_source_and_target: bool = False) -> Dict:
    rows = d_in * (2 if use_both_source_and_target else 1)
    return {"edge_weights": [glorot_uniform(np.random.default_rng(seed + l), rows, d_out) for l in range(L)]}


def ggnn_weights(L: int, d: int, seed: int = 2, cell: str = "gru", random_bias: bool = False) -> Dict:
    rng = np.random.default_rng(seed + 1000)
    gates = 3 if cell.lower() == "gru" else 1
    w = rgcn_weights(L, d, d, seed)
    rec = np.concatenate([orthogonal(rng, d, d) for _ in range(gates)], axis=1)
    bias = (0.1 * rng.standard_normal(gates * d)).astype(np.float32) if random_bias else np.zeros(gates * d, np.float32)
    w["cell"] = {"kernel": glorot_uniform(rng, d, gates * d), "recurrent_kernel": rec.astype(np.float32), "bias": bias}
    return w


def rgat_weights(L: int, d_in: int, d_out: int, seed: int = 2) -> Dict:
    w = rgcn_weights(L, d_in, d_out, seed)
    rng = np.random.default_rng(seed + 2000)
    # tf.get_variable(shape=(2*state_dim)) default initialiser: Glorot-uniform over a 1-D shape
    # (fan_in = fan_out = 2*state_dim)
    w["attention"] = [glorot_uniform(rng, 2 * d_out, 2 * d_out, shape=(2 * d_out,)) for _ in range(L)]
    return w


def _ln(rng: Optional[np.random.Generator], T: int, d: int, randomize: bool):
    if not randomize:
        return [np.ones(d, np.float32) for _ in range(T)], [np.zeros(d, np.float32) for _ in range(T)]
    return ([(1.0 + 0.2 * rng.standard_normal(d)).astype(np.float32) for _ in range(T)],
            [(0.2 * rng.standard_normal(d)).astype(np.float32) for _ in range(T)])


def film_weights(L: int, d_in: int, d_out: int, seed: int = 2, num_timesteps: int = 1, random_ln: bool = False) -> Dict:
    w = rgcn_weights(L, d_in, d_out, seed)
    rng = np.random.default_rng(seed + 3000)
    w["film_weights"] = [glorot_uniform(rng, d_in, 2 * d_out) for _ in range(L)]
    w["ln_gamma"], w["ln_beta"] = _ln(rng, num_timesteps, d_out, random_ln)
    return w


def mlp_kernels(rng: np.random.Generator, d_in: int, d_out: int, hidden_layers: int) -> List[np.ndarray]:
    """utils/utils.py:77-126: hidden_layers hidden Dense layers of width out_size, then the output Dense."""
    dims = [d_in] + [d_out] * hidden_layers + [d_out]
    return [glorot_uniform(rng, dims[i], dims[i + 1]) for i in range(len(dims) - 1)]


def edge_mlp_weights(L: int, d_in: int, d_out: int, num_edge_hidden_layers: int = 1, use_target_state_as_input: bool = True,
                     seed: int = 2, num_timesteps: int = 1, random_ln: bool = False) -> Dict:
    rng = np.random.default_rng(seed + 4000)
    rows = d_in * (2 if use_target_state_as_input else 1)
    w = {"edge_mlps": [mlp_kernels(rng, rows, d_out, num_edge_hidden_layers) for _ in range(L)]}
    w["ln_gamma"], w["ln_beta"] = _ln(rng, num_timesteps, d_out, random_ln)
    return w


def rgin_weights(L: int, d_in: int, d_out: int, num_edge_MLP_hidden_layers: Optional[int] = 1,
                 num_aggr_MLP_hidden_layers: Optional[int] = None, use_target_state_as_input: bool = False,
                 seed: int = 2, num_timesteps: int = 1, random_ln: bool = False) -> Dict:
    rng = np.random.default_rng(seed + 5000)
    rows = d_in * (2 if use_target_state_as_input else 1)
    w: Dict = {}
    msg_width = rows
    if num_edge_MLP_hidden_layers is not None:
        w["edge_mlps"] = [mlp_kernels(rng, rows, d_out, num_edge_MLP_hidden_layers) for _ in range(L)]
        msg_width = d_out
    if num_aggr_MLP_hidden_layers is not None:
        w["aggr_mlp"] = mlp_kernels(rng, msg_width, d_out, num_aggr_MLP_hidden_layers)
    w["ln_gamma"], w["ln_beta"] = _ln(rng, num_timesteps, d_out, random_ln)
    return w


def rgdcn_weights(L: int, num_channels: int, channel_dim: int, use_full_state: bool = False, tie_channel_weights: bool = False,
                  seed: int = 2, stddev: Optional[float] = None) -> Dict:
    """gnns/rgdcn.py:97-104: per (edge type, channel) a bias-free Dense [D or K, K*K], truncated-normal initialised with
    stddev 1/K^2 (pass a larger ``stddev`` in tests so that the dynamic kernels are not vanishingly small)."""
    rng = np.random.default_rng(seed + 6000)
    rows = num_channels * channel_dim if use_full_state else channel_dim
    sd = 1.0 / (channel_dim ** 2) if stddev is None else stddev
    def one():
        return np.clip(rng.standard_normal((rows, channel_dim * channel_dim)), -2.0, 2.0).astype(np.float32) * np.float32(sd)
    return {"channel_weights": [[one() for _ in range(1 if tie_channel_weights else num_channels)] for _ in range(L)]}


def to_torch(weights, device):
    """Recursively move a weight dict (numpy arrays) to float32 tensors on ``device``."""
    import torch
    if isinstance(weights, dict):
        return {k: to_torch(v, device) for k, v in weights.items()}
    if isinstance(weights, (list, tuple)):
        return [to_torch(v, device) for v in weights]
    if weights is None:
        return None
    return torch.as_tensor(np.ascontiguousarray(weights), dtype=torch.float32).to(device)
