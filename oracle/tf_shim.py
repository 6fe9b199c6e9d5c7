"""TEST INFRASTRUCTURE ONLY.  An eager, torch-float64-backed stand-in for the TensorFlow-1 ops that the reference's
energy/force graph uses on this path (RawSymFunc.py symmetry functions and electrostatics, TFMolInstanceDirect.py
energy_inference / dipole_inference), so that oracle/ref_py.py can execute those functions unmodified without
TensorFlow.  Every "tensor" is a torch tensor; tf.gradients is torch.autograd.  Only what those functions call exists.

Variables: the reference creates its weights with self._variable_with_weight_decay (supplied by the caller) and its
biases with tf.Variable(tf.zeros(...), name='biases'); a caller can queue values for the latter with push_biases()."""
import contextlib

import numpy as np
import torch

float64, float32, int64, int32, bool = torch.float64, torch.float32, torch.int64, torch.int32, torch.bool   # noqa: A001

_bias_queue = []


def push_biases(values):
    _bias_queue.extend(values)


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, (list, tuple)) and any(isinstance(v, torch.Tensor) and v.numel() > 1 for v in x):
        return torch.stack(list(x))                      # tf.gradients() returns a list; TF packs it into one tensor (loss_op)
    if isinstance(x, (list, tuple)) and any(isinstance(v, torch.Tensor) for v in x):
        x = [int(v) if isinstance(v, torch.Tensor) else v for v in x]
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def _ints(shp):
    if isinstance(shp, torch.Tensor):
        return [int(v) for v in shp.reshape(-1)]
    return [int(v) for v in shp]


class _Shape(tuple):
    def as_list(self):
        return list(self)


torch.Tensor.get_shape = lambda self: _Shape(self.shape)   # tensor.get_shape().as_list() (RawSymFunc.py:2250)


def shape(x):
    return _Shape(_t(x).shape)


def cast(x, dtype=None, name=None):
    return _t(x).to(dtype)


def constant(v, dtype=None):
    return _t(v, dtype)


def convert_to_tensor(v, dtype=None):
    return _t(v, dtype)


def identity(x, name=None):
    return x


_weight_queue = []


def push_weights(values):
    """Values for the variables TFInstance._variable_with_weight_decay creates (tf.Variable(tf.truncated_normal(...)))."""
    _weight_queue.extend(values)


def truncated_normal(shp, stddev=1.0, dtype=None):
    W = _weight_queue.pop(0)
    assert list(W.shape) == _ints(shp), (W.shape, shp)
    return W


def Variable(init, trainable=True, dtype=None, name=None):
    if name is not None and name.startswith("biases") and _bias_queue:   # 'biases' / 'biaseslayer<i>' (TFMolInstanceDirect.py:5189)
        return _t(_bias_queue.pop(0), float64)
    return _t(init, dtype)


def zeros(shp, dtype=float64):
    return torch.zeros(_ints(shp), dtype=dtype)


def zeros_like(x, dtype=None):
    return torch.zeros_like(_t(x), dtype=dtype)


def ones_like(x, dtype=None):
    return torch.ones_like(_t(x), dtype=dtype)


def reshape(x, shp, name=None):
    return _t(x).reshape(_ints(shp))


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def tile(x, reps):
    return _t(x).repeat(*_ints(reps))


def transpose(x, perm):
    return _t(x).permute(*perm)


def concat(xs, axis, name=None):
    return torch.cat([_t(a) for a in xs], dim=axis)


def stack(xs, axis=0):
    return torch.stack([_t(a) for a in xs], dim=axis)


def range(n, dtype=int32):   # noqa: A001
    return torch.arange(int(n), dtype=dtype)


def slice(x, begin, size, name=None):   # noqa: A001
    x = _t(x)
    idx = []
    for b, s in zip(_ints(begin), _ints(size)):
        idx.append(np.s_[b:] if s < 0 else np.s_[b:b + s])
    return x[tuple(idx)]


def sqrt(x):
    return torch.sqrt(_t(x))


def exp(x):
    return torch.exp(_t(x))


def log(x):
    return torch.log(_t(x))


def cos(x):
    return torch.cos(_t(x))


def acos(x):
    return torch.acos(_t(x))


def erfc(x):
    return torch.erfc(_t(x))


def pow(x, y):   # noqa: A001
    return torch.pow(_t(x, float64) if not isinstance(x, torch.Tensor) else x, y)


def multiply(a, b, name=None):
    return _t(a) * _t(b)


def add(a, b, name=None):
    return _t(a) + _t(b)


def subtract(a, b, name=None):
    return _t(a) - _t(b)


def div(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return a // b
    return _t(a) / _t(b)


def matmul(a, b):
    return _t(a) @ _t(b)


def reduce_sum(x, axis=None):
    return torch.sum(_t(x)) if axis is None else torch.sum(_t(x), dim=axis)


def reduce_max(x):
    return torch.max(_t(x))


def greater(a, b):
    return _t(a) > b


def greater_equal(a, b):
    return _t(a) >= b


def less_equal(a, b):
    return _t(a) <= b


def equal(a, b, name=None):
    return _t(a) == _t(b)


def is_nan(x):
    return torch.isnan(_t(x))


def where(cond, x=None, y=None):
    if x is None:
        return torch.nonzero(cond)
    return torch.where(cond, _t(x), _t(y))


def boolean_mask(x, mask):
    return _t(x)[mask]


def gather_nd(params, indices):
    params, indices = _t(params), _t(indices).long()
    k = indices.shape[-1]
    return params[tuple(indices[..., i] for i in np.arange(k))]


def scatter_nd(indices, updates, shp):
    indices = _t(indices).long()
    out = torch.zeros(_ints(shp), dtype=_t(updates).dtype)
    return out.index_put(tuple(indices[:, i] for i in np.arange(indices.shape[1])), _t(updates), accumulate=True)


class SparseTensor:
    def __init__(self, indices, values, dense_shape):
        self.indices, self.values, self.dense_shape = _t(indices).long(), _t(values), _ints(dense_shape)


def sparse_reduce_sum(sp, axis):
    assert axis == 1
    return torch.zeros(sp.dense_shape[0], dtype=sp.values.dtype).index_add(0, sp.indices[:, 0], sp.values)


def cond(pred, f1, f2):
    return f1() if builtins_bool(pred) else f2()


def builtins_bool(v):
    return (v.item() if isinstance(v, torch.Tensor) else v) not in (0, False)


def verify_tensor_all_finite(x, msg):
    return x


_collections = {}


def reset_collections():
    _collections.clear()


def add_to_collection(key, value):
    _collections.setdefault(key, []).append(value)


def get_collection(key, scope=None):
    """Only the 'losses' collection is kept (the loss ops and the weight-decay terms add to it, TFInstance.py:292-295,
    TFMolInstanceDirect.py:4872-4873); variable collections are the caller's own bookkeeping."""
    return list(_collections.get(key, [])) if key == "losses" else []


def add_n(xs, name=None):
    out = xs[0]
    for x in xs[1:]:
        out = out + x
    return out


class GraphKeys:
    TRAINABLE_VARIABLES = "trainable_variables"


@contextlib.contextmanager
def name_scope(name):
    yield


class nn:   # noqa: N801
    @staticmethod
    def dropout(x, keep_prob):
        kp = float(keep_prob)
        assert kp == 1.0, "evaluation runs with keep_prob = 1"
        return x

    @staticmethod
    def l2_loss(x, name=None):
        return (_t(x) ** 2).sum() / 2


CREATE_GRAPH = False      # True: the coordinate gradient stays differentiable (the force term of the training loss)


def gradients(y, x, name=None):
    return list(torch.autograd.grad(y.sum(), x, retain_graph=True, create_graph=CREATE_GRAPH))
