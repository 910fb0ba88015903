"""A numpy-backed, eager stand-in for the few TensorFlow 1.14 ops that LIU-YUXI/SA-GNN's
propagation path touches (TEST INFRASTRUCTURE ONLY -- used by make_golden_model.py in the build
container to EXECUTE the reference's own model.py text; never imported by the product).

Why: TensorFlow 1.14 has no wheel for this interpreter, so model.py cannot be imported as is.
Installing this module under the name ``tensorflow`` lets ``import model`` succeed unmodified and
lets ``Recommender.messagePropagate`` / ``edgeDropout`` and the loop of ``ours()`` (model.py:80-102,
118-134) run with their wiring, argument roles, Jacobi update order, pad / lookup trick and layer
sum exactly as the reference wrote them.  What this file states on its own is only the per-op
semantics of the TF CPU kernels (each function says which):

  tf.slice, tf.squeeze, tf.range, tf.cast, tf.pad, tf.maximum, tf.add_n, tf.stack, tf.transpose,
  tf.nn.embedding_lookup (GatherV2 axis 0, out-of-range ids raise like the CPU kernel),
  tf.math.segment_sum (rank-1 sorted ids or raise; ids[-1]+1 output rows; gaps are zero rows),
  tf.nn.dropout (TF 1.14 positional keep_prob), tf.sparse.SparseTensor, tf.get_variable,
  tensorflow.contrib.layers.xavier_initializer.

For the rows AFTER the path (interval fusion model.py:135-155 with Utils/attention.py:31-78,
the prediction / SSL pair scores and the meta-weight hinge of model.py:170-173,175-201), used by
make_golden_downstream.py, a second family of stand-ins (same rule: wiring from the reference text, per-op
semantics stated here):

  tf.contrib.rnn.BasicLSTMCell / DropoutWrapper / MultiRNNCell, tf.nn.dynamic_rnn, tf.layers.dense,
  tf.contrib.layers.layer_norm, tf.matmul, tf.exp, tf.tanh, tf.reshape, tf.expand_dims, tf.tile, tf.concat,
  tf.reduce_sum / reduce_mean, tf.shape, tf.stop_gradient, tf.nn.sigmoid / softmax, tf.random_uniform, tf.name_scope.

Variables those create go through ``VarStore`` (creation order + names), so the generator can run the text once to
discover them, move them off their trivial initial values and replay the same text with the stored values.

Everything is eager: a ``Tensor`` wraps an ndarray and keeps its dtype (float32 like the
reference, or float64 when the caller feeds float64 parameters for finite differences).
"""
from __future__ import annotations

import sys
import types

import numpy as np


class InvalidArgumentError(ValueError):
    """What TF raises from shape inference / CPU kernels on bad arguments."""


class Tensor:
    __array_priority__ = 1000

    def __init__(self, a):
        self.a = a.a if isinstance(a, Tensor) else np.asarray(a)

    # -- what NNLayers.FC / model.py use --------------------------------------------------
    def get_shape(self):
        return [int(s) for s in self.a.shape]

    @property
    def shape(self):
        return tuple(self.a.shape)

    @property
    def dtype(self):
        return self.a.dtype

    def __getitem__(self, k):
        return Tensor(self.a[k])

    def _b(self, o):
        if isinstance(o, Tensor):
            return o.a
        if np.ndim(o) == 0 and np.issubdtype(self.a.dtype, np.floating):
            return self.a.dtype.type(o)      # TF converts a Python / numpy scalar operand to the tensor's dtype
        return o

    def __add__(self, o): return Tensor(self.a + self._b(o))
    def __radd__(self, o): return Tensor(self._b(o) + self.a)
    def __sub__(self, o): return Tensor(self.a - self._b(o))
    def __rsub__(self, o): return Tensor(self._b(o) - self.a)
    def __mul__(self, o): return Tensor(self.a * self._b(o))
    def __rmul__(self, o): return Tensor(self._b(o) * self.a)
    def __truediv__(self, o): return Tensor(self.a / self._b(o))
    def __floordiv__(self, o): return Tensor(self.a // self._b(o))
    def __neg__(self): return Tensor(-self.a)
    def __matmul__(self, o): return Tensor(self.a @ self._b(o))
    def __repr__(self): return f"ShimTensor(shape={self.a.shape}, dtype={self.a.dtype})"


def _a(x):
    return x.a if isinstance(x, Tensor) else np.asarray(x)


class Stats:
    """Side channel for the golden generator: how close any LeakyReLU argument came to its kink."""
    min_gap = np.inf

    @classmethod
    def reset(cls):
        cls.min_gap = np.inf


# ---- ops -------------------------------------------------------------------------------------
def tf_slice(x, begin, size):
    """tf.slice: size -1 = everything from ``begin`` to the end of that axis."""
    x = _a(x)
    begin, size = [int(_a(b)) for b in begin], [int(_a(s)) for s in size]      # sizes may be shape-derived tensors
    idx = tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return Tensor(x[idx])


def squeeze(x, axis=None):
    """tf.squeeze: drops every size-1 axis (a [1,1] tensor becomes a scalar)."""
    return Tensor(np.squeeze(_a(x), axis=axis))


def tf_range(*args, dtype=None):
    return Tensor(np.arange(*args, dtype=dtype or np.int32))


def cast(x, dtype):
    """tf.cast: float -> int truncates toward zero (C cast), like the Eigen kernel."""
    return Tensor(_a(x).astype(dtype))


def pad(x, paddings):
    return Tensor(np.pad(_a(x), paddings))


def maximum(a, b):
    a, b = _a(a), _a(b)
    if np.ndim(a) and np.ndim(b) and a.shape == b.shape and a.size:
        g = np.abs(a - b)
        g = g[g > 0]          # exact zeros are rows without edges (constant, on neither side of the kink)
        if g.size:
            Stats.min_gap = min(Stats.min_gap, float(g.min()))
    return Tensor(np.maximum(a, b))


def add_n(ts):
    """tf.add_n: inputs[0] + inputs[1] + ... in list order (AddN CPU kernel)."""
    out = _a(ts[0]).copy()
    for t in ts[1:]:
        out = out + _a(t)
    return Tensor(out)


def stack(ts, axis=0):
    return Tensor(np.stack([_a(t) for t in ts], axis=axis))


def transpose(x, perm=None):
    return Tensor(np.transpose(_a(x), perm))


def embedding_lookup(params, ids):
    """tf.nn.embedding_lookup on one shard = GatherV2(axis=0); the CPU kernel rejects ids outside
    [0, rows)."""
    p, i = _a(params), _a(ids)
    if i.size and (i.min() < 0 or i.max() >= p.shape[0]):
        raise InvalidArgumentError(f"indices = {int(i.max())} is not in [0, {p.shape[0]})")
    return Tensor(p[i])


def segment_sum(data, segment_ids):
    """tf.math.segment_sum: shape inference requires rank-1 ids of data's leading size; the CPU
    kernel requires non-decreasing ids starting at >= 0, emits ids[-1]+1 rows, leaves rows of
    skipped ids zero and adds the members of a segment in their stored order."""
    d, s = _a(data), _a(segment_ids)
    if s.ndim != 1:
        raise InvalidArgumentError(f"Shape must be rank 1 but is rank {s.ndim} (segment_ids)")
    if d.ndim < 1 or d.shape[0] != s.shape[0]:
        raise InvalidArgumentError("segment_ids should be the same size as dimension 0 of input")
    if s.size == 0:
        return Tensor(np.zeros((0,) + d.shape[1:], d.dtype))
    if s[0] < 0 or np.any(np.diff(s) < 0):
        raise InvalidArgumentError("segment ids are not increasing")
    out = np.zeros((int(s[-1]) + 1,) + d.shape[1:], d.dtype)
    np.add.at(out, s, d)                 # unbuffered: members added one by one in edge order, like the CPU kernel
    return Tensor(out)


_drop_rng = np.random.default_rng(12345)


def dropout(x, keep_prob=None, noise_shape=None, seed=None, name=None, rate=None):
    """tf.nn.dropout, TF 1.14 signature: second positional argument is keep_prob."""
    x = _a(x)
    if rate is not None:
        keep_prob = 1.0 - rate
    kp = float(_a(keep_prob))
    keep = _drop_rng.random(x.shape) < kp
    return Tensor(np.where(keep, x / kp, 0).astype(x.dtype))


class SparseTensor:
    """tf.sparse.SparseTensor(indices, values, dense_shape): indices become int64 [E, 2]."""

    def __init__(self, indices, values, dense_shape):
        self.indices = Tensor(np.asarray(_a(indices), dtype=np.int64))
        self.values = Tensor(_a(values))
        self.dense_shape = Tensor(np.asarray(_a(dense_shape), dtype=np.int64))


_init_rng = np.random.default_rng(2024)


def xavier_initializer(uniform=True, seed=None, dtype=np.float32):
    def init(shape):
        shape = [int(s) for s in (shape if np.ndim(shape) else [shape])]
        fan_in = shape[-2] if len(shape) > 1 else shape[-1]
        fan_out = shape[-1]
        for s in shape[:-2]:
            fan_in *= s
            fan_out *= s
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return _init_rng.uniform(-lim, lim, size=shape).astype(np.float32)
    return init


def get_variable(name=None, shape=None, dtype=None, initializer=None, trainable=True, **kw):
    """tf.get_variable as Utils/NNLayers.defineParam calls it (an initializer callable + shape, or an initial
    value); recorded in the VarStore defined below under its own name."""
    return VarStore.get(name, (lambda: initializer(shape)) if callable(initializer) else (lambda: _a(initializer)))



# ---- variables of the consumer rows (dense / LSTM / layer-norm weights) --------------------------
class VarStore:
    """Creation-ordered variable store.  mode "create": every request makes a fresh variable from its TF 1.14
    initializer and records (name, value); mode "replay": the k-th request returns the k-th recorded value (the
    generator perturbs them in between so that biases / gamma / beta are not at 0 / 1)."""
    mode = "create"
    names: list = []
    values: list = []
    cursor = 0

    @classmethod
    def reset(cls):
        cls.mode, cls.names, cls.values, cls.cursor = "create", [], [], 0

    @classmethod
    def replay(cls, dtype=None):
        cls.mode, cls.cursor = "replay", 0
        if dtype is not None:
            cls.values = [v.astype(dtype) for v in cls.values]

    @classmethod
    def get(cls, name, make):
        if cls.mode == "replay":
            assert cls.names[cls.cursor].split("#")[0] == name, (cls.names[cls.cursor], name)
            v = cls.values[cls.cursor]
            cls.cursor += 1
            return Tensor(v)
        v = np.asarray(make())
        cls.names.append("%s#%d" % (name, len(cls.names)))
        cls.values.append(v)
        return Tensor(v)


def _glorot_uniform(shape):
    """The default initializer of tf.get_variable / tf.layers.dense / LSTM kernels in TF 1.14."""
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return _init_rng.uniform(-lim, lim, size=shape).astype(np.float32)


def dense(inputs, units, kernel_initializer=None, use_bias=True, activation=None, name=None):
    """tf.layers.dense: a NEW kernel [in, units] + zero-initialised bias [units] per call (no reuse), applied
    to the last axis; no activation unless given."""
    x = _a(inputs)
    k = VarStore.get("dense/kernel", lambda: (kernel_initializer or _glorot_uniform)([x.shape[-1], units]))
    y = x @ _a(k).astype(x.dtype)
    if use_bias:
        y = y + _a(VarStore.get("dense/bias", lambda: np.zeros(units, np.float32))).astype(x.dtype)
    return Tensor(y if activation is None else _a(activation(Tensor(y))))


def layer_norm(inputs, center=True, scale=True, begin_norm_axis=1, begin_params_axis=-1):
    """tf.contrib.layers.layer_norm (TF 1.14): moments over axes begin_norm_axis..rank-1 (default: every axis
    but the batch axis -- for a [R,T,d] input that is T AND d), beta (zeros) / gamma (ones) of the last axis'
    shape, tf.nn.batch_normalization with variance_epsilon 1e-12."""
    x = _a(inputs)
    axes = tuple(range(begin_norm_axis % x.ndim, x.ndim))
    pshape = x.shape[begin_params_axis:]
    beta = _a(VarStore.get("LayerNorm/beta", lambda: np.zeros(pshape, np.float32))).astype(x.dtype)
    gamma = _a(VarStore.get("LayerNorm/gamma", lambda: np.ones(pshape, np.float32))).astype(x.dtype)
    mean = x.mean(axis=axes, keepdims=True)
    var = np.mean(np.square(x - mean), axis=axes, keepdims=True)
    inv = gamma / np.sqrt(var + 1e-12)          # batch_normalization: inv = rsqrt(var + eps) * scale
    return Tensor(x * inv + (beta - mean * inv))


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


class BasicLSTMCell:
    """tf.contrib.rnn.BasicLSTMCell(num_units): forget_bias 1.0, tanh, state (c, h); kernel
    [input + num_units, 4 * num_units] applied to concat([inputs, h], 1), bias zeros; gates split in the order
    i, j, f, o; new_c = c * sigmoid(f + forget_bias) + sigmoid(i) * tanh(j); new_h = tanh(new_c) * sigmoid(o).
    A Layer builds its variables ONCE: a second dynamic_rnn over the same cell object reuses them."""

    def __init__(self, num_units, forget_bias=1.0):
        self.n, self.forget_bias, self.kernel, self.bias = int(num_units), forget_bias, None, None

    def zero_state(self, batch, dtype):
        return (np.zeros((batch, self.n), dtype), np.zeros((batch, self.n), dtype))

    def __call__(self, x, state):
        c, h = state
        if self.kernel is None:
            self.kernel = VarStore.get("basic_lstm_cell/kernel", lambda: _glorot_uniform([x.shape[1] + self.n, 4 * self.n]))
            self.bias = VarStore.get("basic_lstm_cell/bias", lambda: np.zeros(4 * self.n, np.float32))
        g = np.concatenate([x, h], axis=1) @ _a(self.kernel).astype(x.dtype) + _a(self.bias).astype(x.dtype)
        i, j, f, o = np.split(g, 4, axis=1)
        new_c = c * _sigmoid(f + self.forget_bias) + _sigmoid(i) * np.tanh(j)
        new_h = np.tanh(new_c) * _sigmoid(o)
        return new_h, (new_c, new_h)


class DropoutWrapper:
    """tf.contrib.rnn.DropoutWrapper(cell, output_keep_prob): dropout on the cell OUTPUT only (not the state)."""

    def __init__(self, cell, input_keep_prob=1.0, output_keep_prob=1.0, state_keep_prob=1.0):
        assert input_keep_prob == 1.0 and state_keep_prob == 1.0
        self.cell, self.keep = cell, output_keep_prob

    def zero_state(self, batch, dtype):
        return self.cell.zero_state(batch, dtype)

    def __call__(self, x, state):
        out, st = self.cell(x, state)
        if float(_a(self.keep)) < 1.0:
            out = _a(dropout(out, self.keep))
        return out, st


class MultiRNNCell:
    def __init__(self, cells, state_is_tuple=True):
        assert state_is_tuple
        self.cells = list(cells)

    def zero_state(self, batch, dtype):
        return tuple(c.zero_state(batch, dtype) for c in self.cells)

    def __call__(self, x, state):
        new = []
        for c, s in zip(self.cells, state):
            x, s2 = c(x, s)
            new.append(s2)
        return x, tuple(new)


def dynamic_rnn(cell, inputs, dtype=None, time_major=False, sequence_length=None, initial_state=None):
    """tf.nn.dynamic_rnn, batch-major [B, T, in]: zero initial state, steps t = 0..T-1, outputs stacked on axis 1."""
    assert not time_major and sequence_length is None and initial_state is None
    x = _a(inputs)
    state = cell.zero_state(x.shape[0], x.dtype)
    outs = []
    for t in range(x.shape[1]):
        o, state = cell(x[:, t], state)
        outs.append(o)
    return Tensor(np.stack(outs, axis=1)), state


def matmul(a, b):
    return Tensor(np.matmul(_a(a), _a(b)))


def reshape(x, shape):
    return Tensor(np.reshape(_a(x), [int(_a(s)) for s in shape]))


def concat(ts, axis):
    return Tensor(np.concatenate([_a(t) for t in ts], axis=axis))


def reduce_sum(x, axis=None, keepdims=False):
    return Tensor(np.sum(_a(x), axis=axis, keepdims=keepdims))


def reduce_mean(x, axis=None, keepdims=False):
    return Tensor(np.mean(_a(x), axis=axis, keepdims=keepdims))


def softmax(x, axis=-1):
    x = _a(x)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return Tensor(e / e.sum(axis=axis, keepdims=True))


class _Scope:
    def __init__(self, *a, **k): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"tf1_shim: {name} is outside the propagation path")
    return f


def install():
    """Registers the stand-in under the module names model.py / Utils import."""
    tf = types.ModuleType("tensorflow")
    tf.__dict__.update(
        float32=np.float32, float64=np.float64, int32=np.int32, int64=np.int64, Tensor=Tensor,
        slice=tf_slice, squeeze=squeeze, range=tf_range, cast=cast, pad=pad, maximum=maximum,
        add_n=add_n, stack=stack, transpose=transpose, get_variable=get_variable,
        zeros=lambda shape, dtype=np.float32: Tensor(np.zeros(shape, dtype)),
        ones=lambda shape, dtype=np.float32: Tensor(np.ones(shape, dtype)),
        placeholder=_unsupported("placeholder"), Variable=_unsupported("Variable"),
        matmul=matmul, reshape=reshape, concat=concat, reduce_sum=reduce_sum, reduce_mean=reduce_mean,
        exp=lambda x: Tensor(np.exp(_a(x))), tanh=lambda x: Tensor(np.tanh(_a(x))),
        expand_dims=lambda x, axis: Tensor(np.expand_dims(_a(x), axis)),
        tile=lambda x, m: Tensor(np.tile(_a(x), [int(_a(v)) for v in m])),
        shape=lambda x: Tensor(np.asarray(_a(x).shape, np.int32)),
        stop_gradient=lambda x: Tensor(_a(x)),
        random_uniform=lambda shape, minval=0.0, maxval=1.0, dtype=np.float32:
            Tensor(_init_rng.uniform(minval, maxval, size=shape).astype(np.float32)),
        name_scope=_Scope, variable_scope=_Scope,
    )
    nn = types.ModuleType("tensorflow.nn")
    nn.embedding_lookup = embedding_lookup
    nn.dropout = dropout
    nn.dynamic_rnn = dynamic_rnn
    nn.sigmoid = lambda x: Tensor(_sigmoid(_a(x)))
    nn.tanh = lambda x: Tensor(np.tanh(_a(x)))
    nn.softmax = softmax
    tf_layers = types.ModuleType("tensorflow.layers")
    tf_layers.dense = dense
    tf.layers = tf_layers
    math = types.ModuleType("tensorflow.math")
    math.segment_sum = segment_sum
    sparse = types.ModuleType("tensorflow.sparse")
    sparse.SparseTensor = SparseTensor
    tf.nn, tf.math, tf.sparse = nn, math, sparse
    tf.errors = types.SimpleNamespace(InvalidArgumentError=InvalidArgumentError)
    contrib = types.ModuleType("tensorflow.contrib")
    layers = types.ModuleType("tensorflow.contrib.layers")
    layers.xavier_initializer = xavier_initializer
    layers.layer_norm = layer_norm
    contrib.layers = layers
    rnn = types.ModuleType("tensorflow.contrib.rnn")
    rnn.BasicLSTMCell, rnn.DropoutWrapper, rnn.MultiRNNCell = BasicLSTMCell, DropoutWrapper, MultiRNNCell
    contrib.rnn = rnn
    tf.contrib = contrib
    core = types.ModuleType("tensorflow.core")
    protobuf = types.ModuleType("tensorflow.core.protobuf")
    config_pb2 = types.ModuleType("tensorflow.core.protobuf.config_pb2")
    protobuf.config_pb2 = config_pb2
    core.protobuf = protobuf
    tf.core = core
    mods = {"tensorflow": tf, "tensorflow.nn": nn, "tensorflow.math": math, "tensorflow.sparse": sparse,
            "tensorflow.contrib": contrib, "tensorflow.contrib.layers": layers,
            "tensorflow.contrib.rnn": rnn, "tensorflow.layers": tf_layers, "tensorflow.core": core,
            "tensorflow.core.protobuf": protobuf, "tensorflow.core.protobuf.config_pb2": config_pb2}
    # model.py:4 imports a matplotlib helper it never uses; matplotlib is not installed here
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.cbook  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            cbook = types.ModuleType("matplotlib.cbook")
            cbook.silent_list = list
            mpl.cbook = cbook
            mods.update({"matplotlib": mpl, "matplotlib.cbook": cbook})
    sys.modules.update(mods)
    return tf
