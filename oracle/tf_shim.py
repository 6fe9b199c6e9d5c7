"""TEST INFRASTRUCTURE ONLY.  A numpy stand-in for the handful of TensorFlow-1 ops that the reference's electrostatics
functions use (RawSymFunc.py: DifferenceVectorsLinear, AllDoublesSet, TFCoulombEluSRDSFLR, TFVdwPolyLR,
TFVdwPolyLRWithEle), so that oracle/ref_py.py can execute those functions unmodified, in float64, without TensorFlow.
Eager semantics: every "tensor" is a numpy array.  Only what those functions call is implemented."""
import numpy as np
from scipy.special import erfc as _erfc

float64, float32, int64, int32 = np.float64, np.float32, np.int64, np.int32


def shape(x):
    return np.array(np.shape(x), dtype=np.int64)


def cast(x, dtype):
    return np.asarray(x).astype(dtype)


def constant(v, dtype=None):
    return np.asarray(v, dtype=dtype)


def reshape(x, shp):
    return np.reshape(x, [int(s) for s in np.asarray(shp).ravel()] if not isinstance(shp, (list, tuple)) else [int(s) for s in shp])


def sqrt(x):
    return np.sqrt(x)


def exp(x):
    return np.exp(x)


def log(x):
    return np.log(x)


def erfc(x):
    return _erfc(x)


def pow(x, y):   # noqa: A001
    return np.power(x, y)


def multiply(a, b):
    return np.multiply(a, b)


def reduce_sum(x, axis=None):
    return np.sum(x, axis=axis)


def greater(a, b):
    return np.greater(a, b)


def equal(a, b):
    return np.equal(a, b)


def is_nan(x):
    return np.isnan(x)


def zeros_like(x):
    return np.zeros_like(x)


def ones_like(x):
    return np.ones_like(x)


def where(cond, x=None, y=None):
    if x is None:
        return np.argwhere(cond).astype(np.int64)
    return np.where(cond, x, y)


def range(n, dtype=np.int32):   # noqa: A001
    return np.arange(int(n), dtype=dtype)


def slice(x, begin, size):   # noqa: A001
    x = np.asarray(x)
    idx = []
    for d, (b, s) in enumerate(zip(begin, size)):
        b, s = int(b), int(s)
        idx.append(np.s_[b:] if s < 0 else np.s_[b:b + s])
    return x[tuple(idx)]


def concat(xs, axis):
    return np.concatenate([np.asarray(a) for a in xs], axis=axis)


def stack(xs, axis=0):
    return np.stack([np.asarray(a) for a in xs], axis=axis)


def tile(x, reps):
    return np.tile(x, [int(r) for r in reps])


def transpose(x, perm):
    return np.transpose(x, perm)


def gather_nd(params, indices):
    params, indices = np.asarray(params), np.asarray(indices).astype(np.int64)
    k = indices.shape[-1]
    return params[tuple(indices[..., i] for i in np.arange(k))]


class SparseTensor:
    def __init__(self, indices, values, dense_shape):
        self.indices, self.values, self.dense_shape = np.asarray(indices), np.asarray(values), [int(s) for s in dense_shape]


def sparse_reduce_sum(sp, axis):
    assert axis == 1
    return np.bincount(sp.indices[:, 0], weights=sp.values, minlength=sp.dense_shape[0])
