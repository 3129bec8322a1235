"""CPU: the numpy stand-in for TensorFlow 1.13 / dpu_utils (tests/tf1_shim) -- every KERNEL semantic it restates, as an executable
statement with hand-computed values.  These are the assumptions that remain after the reference's own code has been executed
(SURVEY.md Appendix A); tests/golden/make_tf1_fixtures.py checks them against a real TensorFlow where one exists."""
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tf1_shim                               # noqa: E402


@pytest.fixture()
def tf():
    session = tf1_shim.Session(np.float64, seed=0)
    tf1_shim.build_modules(session)
    session.tf._session = session
    return session.tf


def test_unsorted_segment_ops(tf):
    data = np.array([[1.0, -2.0], [3.0, 5.0], [10.0, 0.5]])
    ids = np.array([2, 0, 2])
    np.testing.assert_array_equal(tf.unsorted_segment_sum(data, ids, 4), [[3, 5], [0, 0], [11, -1.5], [0, 0]])
    np.testing.assert_array_equal(tf.unsorted_segment_mean(data, ids, 4), [[3, 5], [0, 0], [5.5, -0.75], [0, 0]])       # empty: 0 / max(n, 1)
    np.testing.assert_allclose(tf.unsorted_segment_sqrt_n(data, ids, 4)[2], np.array([11, -1.5]) / math.sqrt(2))
    mx = tf.unsorted_segment_max(data, ids, 4)
    np.testing.assert_array_equal(mx[[0, 2]], [[3, 5], [10, 0.5]])
    assert mx[1, 0] == float(np.finfo(np.float32).min)                       # empty segment: numeric_limits<float>::lowest()


def test_activations_and_dropout(tf):
    x = np.array([-2.0, -0.5, 0.0, 0.5, 2.0])
    np.testing.assert_allclose(tf.nn.leaky_relu(x), [-0.4, -0.1, 0.0, 0.5, 2.0])                                         # alpha = 0.2
    np.testing.assert_allclose(tf.nn.elu(x), [math.expm1(-2.0), math.expm1(-0.5), 0.0, 0.5, 2.0])
    np.testing.assert_allclose(tf.nn.selu(x)[0], 1.0507009873554805 * 1.6732632423543772 * math.expm1(-2.0))
    np.testing.assert_allclose(tf.nn.selu(x)[-1], 1.0507009873554805 * 2.0)
    np.testing.assert_allclose(tf.erf(x / tf.sqrt(2.0)), [math.erf(v / math.sqrt(2.0)) for v in x])
    np.testing.assert_array_equal(tf.nn.dropout(x, rate=0.0), x)
    with pytest.raises(AssertionError):
        tf.nn.dropout(x, rate=0.1)                                            # only the evaluation path exists here


def test_dense_layers_create_variables_under_tf_names(tf):
    s = tf._session
    with tf.variable_scope("graph_model"), tf.variable_scope("gnn_layer_0"):
        d = tf.keras.layers.Dense(units=3, use_bias=False, activation=None, name="Edge_0_Weight")
        y = d(np.ones((2, 4)))
        with tf.variable_scope("Edge_1_MLP"):
            a, b = tf.layers.Dense(units=5, use_bias=False, activation=tf.nn.relu), tf.layers.Dense(units=2, use_bias=False, activation=None)
            z = b(a(np.ones((2, 4))))
        v = tf.get_variable(shape=(6), name="Edge_0_Attention_Parameters")
    assert y.shape == (2, 3) and z.shape == (2, 2) and v.shape == (6,)
    assert list(s.variables) == ["graph_model/gnn_layer_0/Edge_0_Weight/kernel:0", "graph_model/gnn_layer_0/Edge_1_MLP/dense/kernel:0",
                                 "graph_model/gnn_layer_0/Edge_1_MLP/dense_1/kernel:0", "graph_model/gnn_layer_0/Edge_0_Attention_Parameters:0"]
    k = s.variables["graph_model/gnn_layer_0/Edge_0_Weight/kernel:0"]
    assert np.all(np.abs(k) <= math.sqrt(6.0 / (4 + 3)))                      # Glorot-uniform limit
    np.testing.assert_allclose(y, np.ones((2, 4)) @ k)                        # y = x . kernel, kernel [in, out]


def test_gru_cell_keras_tf113_defaults(tf):
    """hard_sigmoid recurrent activation, reset_after = False, gate order z | r | h, h' = z*h + (1-z)*hh."""
    cell = tf.keras.layers.GRUCell(2, activation=np.tanh)
    x, h = np.array([[0.3, -0.7]]), np.array([[0.5, 0.1]])
    out, (state,) = cell(x, [h])
    s = tf._session
    k, u, b = (s.variables["gru_cell/%s:0" % n] for n in ("kernel", "recurrent_kernel", "bias"))
    assert k.shape == (2, 6) and u.shape == (2, 6) and b.shape == (6,) and not b.any()
    np.testing.assert_allclose(u[:, :].T @ u[:, :], u.T @ u)                  # (shape only; orthogonality is a property of the initialiser)
    hs = lambda t: np.clip(0.2 * t + 0.5, 0, 1)   # noqa: E731
    z = hs(x @ k[:, :2] + h @ u[:, :2])
    r = hs(x @ k[:, 2:4] + h @ u[:, 2:4])
    hh = np.tanh(x @ k[:, 4:] + (r * h) @ u[:, 4:])
    np.testing.assert_allclose(out, z * h + (1 - z) * hh, rtol=1e-14)
    assert state is out
    with pytest.raises(ValueError):
        tf.keras.layers.LSTMCell(2)(x, [h])                                   # one state passed (gnns/ggnn.py:92): fails like Keras


def test_simple_rnn_cell(tf):
    cell = tf.keras.layers.SimpleRNNCell(3, activation=tf.nn.relu)
    x, h = np.array([[1.0, 2.0]]), np.array([[0.1, 0.2, 0.3]])
    out, _ = cell(x, [h])
    s = tf._session
    k, u = s.variables["simple_rnn_cell/kernel:0"], s.variables["simple_rnn_cell/recurrent_kernel:0"]
    np.testing.assert_allclose(out, np.maximum(x @ k + h @ u, 0), rtol=1e-14)


def test_layer_norm_defaults(tf):
    x = np.array([[1.0, 2.0, 3.0, 6.0], [5.0, 5.0, 5.0, 5.0]])
    y = tf.contrib.layers.layer_norm(x)
    y2 = tf.contrib.layers.layer_norm(x)                                      # a second call opens LayerNorm_1
    mu, var = x[0].mean(), x[0].var()                                         # biased variance
    np.testing.assert_allclose(y[0], (x[0] - mu) / math.sqrt(var + 1e-12), rtol=1e-12)
    np.testing.assert_allclose(y[1], 0.0, atol=1e-5)                          # zero variance: eps 1e-12 keeps it finite
    np.testing.assert_array_equal(y, y2)
    assert [n for n in tf._session.variables] == ["LayerNorm/beta:0", "LayerNorm/gamma:0", "LayerNorm_1/beta:0", "LayerNorm_1/gamma:0"]


def test_unsorted_segment_log_softmax_of_dpu_utils():
    session = tf1_shim.Session(np.float64)
    mods = tf1_shim.build_modules(session)
    f = mods["dpu_utils.tfutils"].unsorted_segment_log_softmax
    logits, ids = np.array([1.0, 2.0, 3.0, -1.0]), np.array([0, 0, 2, 0])
    got = np.exp(f(logits, ids, 3))
    e = np.exp(np.array([1.0, 2.0, -1.0]))
    np.testing.assert_allclose(got[[0, 1, 3]], e / e.sum(), rtol=1e-14)
    np.testing.assert_allclose(got[2], 1.0)


def test_installed_context_leaves_no_trace():
    before = {k for k in sys.modules if k.split(".")[0] in ("tensorflow", "dpu_utils", "gnns", "utils")}
    with tf1_shim.installed(np.float32) as session:
        import tensorflow as shim_tf
        assert getattr(shim_tf, "__shim__", False) and session.tf is shim_tf
    after = {k for k in sys.modules if k.split(".")[0] in ("tensorflow", "dpu_utils", "gnns", "utils")}
    assert after == before


# ---- graph_mode: the eager stand-ins for the scaffold / heads ----
def test_sigmoid_cross_entropy_and_clip_by_norm_against_torch(tf):
    import torch
    rng = np.random.default_rng(0)
    x, z = rng.standard_normal((7, 5)) * 6, (rng.random((7, 5)) < 0.4).astype(np.float64)
    want = torch.nn.functional.binary_cross_entropy_with_logits(torch.as_tensor(x), torch.as_tensor(z), reduction="none").numpy()
    np.testing.assert_allclose(tf.nn.sigmoid_cross_entropy_with_logits(logits=x, labels=z), want, rtol=1e-13, atol=1e-15)
    g = rng.standard_normal((4, 3))
    n = np.linalg.norm(g)
    np.testing.assert_allclose(tf.clip_by_norm(g, n / 2), g / 2)                    # longer than the clamp: rescaled to it
    np.testing.assert_array_equal(tf.clip_by_norm(g, 2 * n), g * (2 * n) / (2 * n))  # shorter: t * clip / max(norm, clip) == t


def test_placeholders_inert_without_feed_and_eager_with_one(tf):
    s = tf._session
    a, b = tf.placeholder(tf.float32, [None, 3], name="x"), tf.placeholder(tf.float32, [None, 3], name="x")
    assert a is not b and len({a: 1, b: 2}) == 2 and a.name == "x"               # feed_dict keys, hashable by identity
    s.feeds = {"x": np.array([[0.1, 0.2, 0.3]]), "n": 7}
    x = tf.placeholder(tf.float32, [None, 3], name="x")
    assert x.dtype == np.float64 and np.array_equal(x, np.float32([[0.1, 0.2, 0.3]]).astype(np.float64))   # what an fp32 placeholder holds
    assert np.array_equal(tf.placeholder(tf.int64, [], name="n"), 7) and tf.placeholder(tf.int64, [], name="n").dtype == np.int64
    assert len({x: 1, tf.placeholder(tf.float32, name="x"): 2}) == 2              # still usable as keys (epoch loop)
    assert float(tf.placeholder_with_default(1.0, [], name="keep")) == 1.0
    with pytest.raises(KeyError):
        tf.placeholder(tf.float32, name="missing")


def test_unnamed_keras_dense_numbered_per_graph_tf_layers_dense_per_scope(tf):
    s = tf._session
    x = np.ones((2, 3))
    with tf.variable_scope("graph_model"):
        tf.keras.layers.Dense(4, use_bias=False)(x)                               # projection (sparse_graph_model.py:166)
        with tf.variable_scope("A"):
            tf.layers.Dense(4, use_bias=False)(x)
            tf.layers.Dense(4, use_bias=False)(x)
        with tf.variable_scope("B"):
            tf.layers.Dense(4, use_bias=False)(x)
    tf.keras.layers.Dense(2)(x)                                                   # PPI head (ppi_task.py:176): the graph's 2nd unnamed Keras layer
    assert list(s.variables) == ["graph_model/dense/kernel:0", "graph_model/A/dense/kernel:0", "graph_model/A/dense_1/kernel:0",
                                 "graph_model/B/dense/kernel:0", "dense_1/kernel:0", "dense_1/bias:0"]


def test_bookkeeping_variable_collections_and_optimizer_stub(tf):
    s = tf._session
    with tf.variable_scope("m"):
        tf.get_variable("w", shape=(3, 2))
    tf.get_variable(name="total_num_graphs", shape=(), dtype=tf.int64, initializer=tf.zeros_initializer, trainable=False)
    assert [v.name for v in tf.trainable_variables()] == ["m/w:0"]
    assert [d.value for d in tf.trainable_variables()[0].get_shape()] == [3, 2]
    g = tf.Graph()
    assert [v.name for v in g.get_collection(tf.GraphKeys.GLOBAL_VARIABLES)] == ["m/w:0", "total_num_graphs:0"]
    sess = tf.Session(graph=g, config=tf.ConfigProto())
    out = sess.run({v.name: v for v in g.get_collection(tf.GraphKeys.GLOBAL_VARIABLES)})
    assert out["total_num_graphs:0"].dtype == np.int64 and out["m/w:0"].shape == (3, 2)
    opt = tf.train.RMSPropOptimizer(learning_rate=0.1, decay=0.5, momentum=0.2)
    gv = opt.compute_gradients(np.float64(1.0), var_list=tf.trainable_variables())
    assert [(g_, v.name) for g_, v in gv] == [(None, "m/w:0")] and s.optimizers == [("RMSPropOptimizer", {"learning_rate": 0.1, "decay": 0.5, "momentum": 0.2})]
    with pytest.raises(ValueError):
        tf.trainable_variables()[0].assign(np.zeros((2, 2)))
