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
        return o.a if isinstance(o, Tensor) else o

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
    if callable(initializer):
        return Tensor(initializer(shape))
    return Tensor(_a(initializer))


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
    )
    nn = types.ModuleType("tensorflow.nn")
    nn.embedding_lookup = embedding_lookup
    nn.dropout = dropout
    math = types.ModuleType("tensorflow.math")
    math.segment_sum = segment_sum
    sparse = types.ModuleType("tensorflow.sparse")
    sparse.SparseTensor = SparseTensor
    tf.nn, tf.math, tf.sparse = nn, math, sparse
    tf.errors = types.SimpleNamespace(InvalidArgumentError=InvalidArgumentError)
    contrib = types.ModuleType("tensorflow.contrib")
    layers = types.ModuleType("tensorflow.contrib.layers")
    layers.xavier_initializer = xavier_initializer
    contrib.layers = layers
    tf.contrib = contrib
    core = types.ModuleType("tensorflow.core")
    protobuf = types.ModuleType("tensorflow.core.protobuf")
    config_pb2 = types.ModuleType("tensorflow.core.protobuf.config_pb2")
    protobuf.config_pb2 = config_pb2
    core.protobuf = protobuf
    tf.core = core
    mods = {"tensorflow": tf, "tensorflow.nn": nn, "tensorflow.math": math, "tensorflow.sparse": sparse,
            "tensorflow.contrib": contrib, "tensorflow.contrib.layers": layers, "tensorflow.core": core,
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
