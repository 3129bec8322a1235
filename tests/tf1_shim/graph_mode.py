"""The graph-building slice of TF1 that the reference's model scaffold and task heads touch (models/sparse_graph_model.py,
models/*_model.py, tasks/{sparse_graph,ppi,qm9}_task.py), as EAGER stand-ins -- TEST INFRASTRUCTURE like the rest of the shim.

The reference builds a static graph over placeholders and later feeds it through ``sess.run``.  Here the feed is known up
front (``session.feeds``: placeholder name -> array), ``tf.placeholder`` hands that array out, and every op of
``Sparse_Graph_Model.__make_model`` runs while the constructor executes; afterwards the model object's private op dictionary
(``_Sparse_Graph_Model__ops``) holds VALUES: 'final_node_representations', 'task_metrics', ...  With ``session.feeds is
None`` placeholders are inert, hashable objects -- what the batchers need for their feed_dict keys.

Only the forward pass exists: optimizers return no gradients (``compute_gradients`` -> (None, var) pairs, which the reference
passes through untouched, sparse_graph_model.py:253-259)."""
import contextlib
import types

import numpy as np

TRAINABLE, GLOBAL = "trainable_variables", "variables"


class _Dim:
    def __init__(self, value):
        self.value = int(value)


class Variable:
    """What tf.trainable_variables() / graph.get_collection() hand out: .name, .get_shape() (dims with .value)."""

    def __init__(self, session, name):
        self._session, self.name = session, name

    def value(self):
        return self._session.variables[self.name]

    def get_shape(self):
        return [_Dim(d) for d in self.value().shape]

    def assign(self, value):
        """variable.assign(value) of load_weights (sparse_graph_model.py:118): shapes must agree, as TF checks."""
        old, new = self.value(), np.asarray(value)
        if old.shape != new.shape:
            raise ValueError("assign to %s: shape %s != %s" % (self.name, new.shape, old.shape))
        self._session.variables[self.name] = new.astype(old.dtype)
        return None

    def __repr__(self):
        return "<variable %s %s>" % (self.name, self.value().shape)


class Placeholder:
    """Inert tf.placeholder (no feed known): a feed_dict key, hashable by identity like a tf.Tensor."""

    def __init__(self, dtype, shape=None, name=None):
        self.dtype, self.shape, self.name = dtype, shape, name

    def __repr__(self):
        return "<placeholder %s>" % self.name


class FeedArray(np.ndarray):
    """What an eager placeholder hands out: the fed array, but hashable BY IDENTITY like a tf.Tensor, because the reference's
    epoch loop uses the model's placeholders as feed_dict keys (sparse_graph_model.py:277-281, tasks/*_task.py batchers)."""

    def __hash__(self):
        return id(self)


def _as_feed(value):
    return np.asarray(value).view(FeedArray)


def _resolve(session, obj):
    """sess.run on an eager world: Variables become their values, containers are walked, arrays pass through."""
    if isinstance(obj, Variable):
        return obj.value()
    if isinstance(obj, dict):
        return {k: _resolve(session, v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_resolve(session, v) for v in obj)
    return obj


def extend(tf, session):
    def placeholder(dtype=None, shape=None, name=None):
        if session.feeds is None:
            return Placeholder(dtype, shape, name)
        if name not in session.feeds:
            raise KeyError("no feed for placeholder %r" % name)
        value = np.asarray(session.feeds[name])
        if dtype is not None and np.issubdtype(np.dtype(dtype), np.floating):
            # the value an fp32 placeholder receives (rounded to float32), carried in the session's float type (fp64 = truth run)
            return _as_feed(value.astype(np.dtype(dtype)).astype(session.dtype))
        return _as_feed(value if dtype is None else value.astype(dtype))

    def placeholder_with_default(default, shape=None, name=None):
        if session.feeds is not None and name in session.feeds:
            return _as_feed(session.dtype(session.feeds[name]))
        return _as_feed(session.dtype(default))

    tf.placeholder, tf.placeholder_with_default = placeholder, placeholder_with_default

    class zeros_initializer:
        pass

    tf.zeros_initializer = zeros_initializer
    base_get_variable = tf.get_variable

    def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
        if initializer is zeros_initializer or isinstance(initializer, zeros_initializer):
            full = (session.scope_path() + "/" if session.scope else "") + name + ":0"
            session.variables[full] = np.zeros(tuple(shape or ()), dtype or session.dtype)
            if not trainable:
                session.non_trainable.add(full)
            return session.variables[full]
        return base_get_variable(name, shape=shape, dtype=dtype, initializer=initializer, trainable=trainable)

    tf.get_variable = get_variable
    tf.assign_add = lambda ref, value, **kw: ref + value
    tf.trainable_variables = lambda: [Variable(session, n) for n in session.variables if n not in session.non_trainable]
    tf.global_variables = lambda: [Variable(session, n) for n in session.variables]

    class GraphKeys:
        TRAINABLE_VARIABLES, GLOBAL_VARIABLES = TRAINABLE, GLOBAL

    class Graph:
        def as_default(self):
            return contextlib.nullcontext(self)

        def get_collection(self, key):
            return tf.trainable_variables() if key == TRAINABLE else tf.global_variables()

    class TfSession:
        def __init__(self, graph=None, config=None):
            self.graph = graph if graph is not None else Graph()

        def run(self, fetches, feed_dict=None):
            """Everything was computed when the graph was built; ``session.run_hook`` (tests of the epoch loop) may script the
            per-batch results instead -- it sees the fetches and the feed_dict the reference assembled."""
            if session.run_hook is not None:
                return session.run_hook(fetches, feed_dict)
            return _resolve(session, fetches)

    class ConfigProto:
        def __init__(self, **kw):
            self.gpu_options = types.SimpleNamespace()

    tf.GraphKeys, tf.Graph, tf.Session, tf.ConfigProto = GraphKeys, Graph, TfSession, ConfigProto
    tf.set_random_seed = lambda seed: None
    tf.name_scope = lambda name=None, *a, **kw: contextlib.nullcontext(name)
    tf.group = lambda *a, **kw: None
    tf.global_variables_initializer = tf.local_variables_initializer = lambda: None
    tf.variables_initializer = lambda var_list, name=None: None

    summary = types.ModuleType("tensorflow.summary")
    summary.scalar = lambda name, tensor, **kw: None
    summary.merge_all = lambda: None
    summary.FileWriter = lambda *a, **kw: None
    tf.summary = summary

    class _Optimizer:
        """tf.train.*Optimizer: records how it was constructed.  No autodiff here: compute_gradients returns (None, var) pairs --
        or what ``session.gradient_hook(variable name, shape)`` prescribes (tests of the clipping step) -- and apply_gradients
        records what it was handed (``session.applied``)."""

        def __init__(self, *a, **kw):
            assert not a, "the reference passes keyword arguments only (sparse_graph_model.py:241-249)"
            self.kwargs = kw
            session.optimizers.append((type(self).__name__, kw))

        def compute_gradients(self, loss, var_list=None):
            session.loss_for_gradients = loss
            hook = session.gradient_hook
            return [(None if hook is None else hook(v.name, v.value().shape), v) for v in (var_list or [])]

        def apply_gradients(self, grads_and_vars, **kw):
            session.applied = [(g, v.name) for g, v in grads_and_vars]
            return None

    train = types.ModuleType("tensorflow.train")
    train.GradientDescentOptimizer = type("GradientDescentOptimizer", (_Optimizer,), {})
    train.RMSPropOptimizer = type("RMSPropOptimizer", (_Optimizer,), {})
    train.AdamOptimizer = type("AdamOptimizer", (_Optimizer,), {})
    tf.train = train

    # ---- the remaining eager ops of the scaffold / heads ----
    tf.zeros_like = lambda x, **kw: np.zeros_like(np.asarray(x))
    tf.squeeze = lambda x, axis=None, **kw: np.squeeze(np.asarray(x), axis=axis)
    tf.abs, tf.square = np.abs, np.square
    tf.reduce_sum = lambda x, axis=None, **kw: np.sum(np.asarray(x), axis=axis)
    tf.reduce_mean = lambda x, axis=None, **kw: np.mean(np.asarray(x), axis=axis)
    tf.constant = lambda value, dtype=None, **kw: np.asarray(value, dtype=dtype)

    def clip_by_norm(t, clip_norm, **kw):
        t = np.asarray(t)
        return t * clip_norm / np.maximum(np.sqrt(np.sum(t * t)), clip_norm)

    tf.clip_by_norm = clip_by_norm

    def sigmoid_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
        """max(x, 0) - x * z + log(1 + exp(-|x|))  (nn_impl.py)."""
        x, z = np.asarray(logits), np.asarray(labels).astype(np.asarray(logits).dtype)
        return np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))

    tf.nn.sigmoid_cross_entropy_with_logits = sigmoid_cross_entropy_with_logits
    return {"tensorflow.summary": summary, "tensorflow.train": train}


class ThreadedIterator:
    """dpu_utils.utils.ThreadedIterator: here simply the wrapped iterator."""

    def __init__(self, original_iterator, max_queue_size=2, enabled=True):
        self._it = original_iterator

    def __iter__(self):
        return iter(self._it)
