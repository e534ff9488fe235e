"""TEST INFRASTRUCTURE ONLY -- a minimal TensorFlow-1.x look-alike backed by torch eager ops.

Purpose: let the *unmodified* reference files under /root/reference/src/Models be imported and
their hot-path functions (prior_kernels, approx_kernels, build_kernels, tf_kernel, gp_vae_sample,
calc_gp_kl, trans_break_mat, gp_kl_div; V2: kernel_matrix, calc_gp_kl, gp_kl_div) be executed
verbatim in the dev container, forward and backward (torch autograd stands in for TF autodiff).
It is used by oracle/gen_golden.py to produce the fixtures in tests/golden/ and by the CPU tests
that pin oracle/gp_kl_oracle.py against the reference.  Nothing in the product path imports it.

Only the ~45 TF1 symbols the hot path touches are provided.  Semantics
that matter and are reproduced here:
  * tf.split(x, n_or_sizes, axis): int -> equal parts, list/tensor -> explicit sizes
  * tf.slice(x, begin, size), tf.pad(x, [[lo,hi],...]) outermost-first
  * tf.constant(v, shape=) broadcasts a scalar and reshapes a list
  * tf.linalg.logdet(M) = 2*sum(log(diag(chol(M))))   (TF computes it through Cholesky)
  * tf.matrix_inverse / tf.matrix_determinant are LU based (torch.linalg.inv / det)
  * tf.random_normal draws through RANDOM_SOURCE so a fixture generator can record the noise
"""
import sys
import types

import torch

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
newaxis = None
_pyslice = slice  # the builtin, before tf.slice shadows it below

# every tf.Variable created while a reference function runs, by name (for gradient read-back)
VARIABLES = {}
# callable(shape:list[int]) -> float32 tensor; replaced by fixture generators
RANDOM_SOURCE = None
RANDOM_LOG = []


def _to_int(v):
    if isinstance(v, torch.Tensor):
        return int(v.reshape(-1)[0].item()) if v.numel() == 1 else [int(a) for a in v.tolist()]
    return int(v)


def _shape_list(shape):
    if isinstance(shape, torch.Tensor):
        return [int(a) for a in shape.reshape(-1).tolist()]
    if isinstance(shape, (tuple, list)):
        return [_to_int(s) for s in shape]
    return [int(shape)]


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    import numpy as np
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        return t if dtype is None else t.to(dtype)
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        return torch.stack([_t(a) for a in x])
    if dtype is None:
        dtype = torch.float32 if isinstance(x, float) or (
            isinstance(x, (list, tuple)) and any(isinstance(a, float) for a in _flatten(x))) else torch.int32
    return torch.tensor(x, dtype=dtype)


def _flatten(x):
    for a in x:
        if isinstance(a, (list, tuple)):
            yield from _flatten(a)
        else:
            yield a


def constant(value, dtype=None, shape=None, name=None):
    t = _t(value, dtype)
    if shape is not None:
        shape = _shape_list(shape)
        n = 1
        for s in shape:
            n *= s
        if t.numel() == 1:
            t = t.reshape(()).expand(shape).clone()
        else:
            flat = t.reshape(-1)
            if flat.numel() < n:  # TF1 pads a too-short list with its last element
                flat = torch.cat([flat, flat[-1:].expand(n - flat.numel())])
            t = flat.reshape(shape)
    return t


def Variable(initial_value, name=None, dtype=None, trainable=True):
    t = _t(initial_value, dtype).detach().clone()
    if t.is_floating_point():
        t.requires_grad_(True)
    VARIABLES[name if name is not None else "var_%d" % len(VARIABLES)] = t
    return t


def identity(x, name=None):
    return x


def placeholder(dtype, shape=None, name=None):
    raise RuntimeError("tf.placeholder: the stub only executes the eager hot-path functions")


def split(value, num_or_size_splits, axis=0, num=None, name=None):
    value = _t(value)
    if isinstance(num_or_size_splits, torch.Tensor):
        sizes = [int(a) for a in num_or_size_splits.reshape(-1).tolist()]
        return list(torch.split(value, sizes, dim=axis))
    if isinstance(num_or_size_splits, (list, tuple)):
        return list(torch.split(value, [_to_int(s) for s in num_or_size_splits], dim=axis))
    n = int(num_or_size_splits)
    assert value.shape[axis] % n == 0, "tf.split: dimension not divisible"
    return list(torch.split(value, value.shape[axis] // n, dim=axis))


def unstack(value, num=None, axis=0, name=None):
    value = _t(value)
    out = list(torch.unbind(value, dim=axis))
    if num is not None:
        assert len(out) == int(num)
    return out


def stack(values, axis=0, name=None):
    return torch.stack([_t(v) for v in values], dim=axis)


def concat(values, axis, name=None):
    return torch.cat([_t(v) for v in values], dim=axis)


def reduce_max(x, axis=None, name=None):
    x = _t(x)
    return x.max() if axis is None else x.max(dim=axis).values


def reduce_sum(x, axis=None, name=None, keepdims=False):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def reduce_mean(x, axis=None, name=None):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=axis)


def pow(x, y, name=None):  # noqa: A001 - mirrors tf.pow
    x = _t(x)
    return torch.pow(x, y)


def square(x, name=None):
    x = _t(x)
    return x * x


def sqrt(x, name=None):
    return torch.sqrt(_t(x))


def exp(x, name=None):
    return torch.exp(_t(x))


def log(x, name=None):
    return torch.log(_t(x))


def sigmoid(x, name=None):
    return torch.sigmoid(_t(x))


def slice(input_, begin, size, name=None):  # noqa: A001 - mirrors tf.slice
    x = _t(input_)
    begin = [_to_int(b) for b in begin]
    size = [_to_int(s) for s in size]
    idx = []
    for d, (b, s) in enumerate(zip(begin, size)):
        e = x.shape[d] if s == -1 else b + s
        idx.append(_pyslice(b, e))
    return x[tuple(idx)]


def reshape(tensor, shape, name=None):
    return _t(tensor).reshape(_shape_list(shape))


def cast(x, dtype, name=None):
    return _t(x).to(dtype)


def pad(tensor, paddings, mode="CONSTANT", name=None, constant_values=0):
    x = _t(tensor)
    assert mode == "CONSTANT"
    flat = []
    for lo, hi in reversed(list(paddings)):  # F.pad wants innermost dimension first
        flat += [_to_int(lo), _to_int(hi)]
    return torch.nn.functional.pad(x, flat, mode="constant", value=constant_values)


def squeeze(x, axis=None, name=None):
    x = _t(x)
    return x.squeeze() if axis is None else x.squeeze(axis)


def eye(num_rows, num_columns=None, dtype=torch.float32, name=None):
    n = _to_int(num_rows)
    return torch.eye(n, n if num_columns is None else _to_int(num_columns), dtype=dtype)


def shape(x, name=None):
    return torch.tensor(list(_t(x).shape), dtype=torch.int32)


def ones(shape, dtype=torch.float32, name=None):
    return torch.ones(_shape_list(shape), dtype=dtype)


def cholesky(x, name=None):
    return torch.linalg.cholesky(_t(x))


def matrix_inverse(x, name=None):
    return torch.linalg.inv(_t(x))


def matrix_determinant(x, name=None):
    return torch.linalg.det(_t(x))


def matmul(a, b, name=None):
    return torch.matmul(_t(a), _t(b))


def transpose(a, perm=None, name=None):
    a = _t(a)
    if perm is None:
        return a.permute(*reversed(range(a.dim())))
    return a.permute(*perm)


def trace(x, name=None):
    return torch.trace(_t(x))


def diag(x, name=None):
    return torch.diag(_t(x))


def scalar_mul(scalar, x, name=None):
    return scalar * _t(x)


def multiply(x, y, name=None):
    return _t(x) * _t(y)


def add(x, y, name=None):
    return _t(x) + _t(y)


def tile(x, multiples, name=None):
    return _t(x).repeat(*[_to_int(m) for m in multiples])


def expand_dims(x, axis, name=None):
    return _t(x).unsqueeze(axis)


def random_normal(shape, mean=0.0, stddev=1.0, dtype=torch.float32, seed=None, name=None):
    shape = _shape_list(shape)
    r = RANDOM_SOURCE(shape) if RANDOM_SOURCE is not None else torch.randn(shape, dtype=dtype)
    RANDOM_LOG.append(r)
    return r * stddev + mean


def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=torch.float32, seed=None, name=None):
    shape = _shape_list(shape)
    return torch.fmod(torch.randn(shape, dtype=dtype), 2.0) * stddev + mean


def _logdet(x, name=None):
    L = torch.linalg.cholesky(_t(x))
    return 2.0 * torch.sum(torch.log(torch.diagonal(L, dim1=-2, dim2=-1)), dim=-1)


def install():
    """Register fake `tensorflow`, `matplotlib(.pyplot)` and `sklearn.externals.joblib` modules."""
    me = sys.modules[__name__]
    tf = types.ModuleType("tensorflow")
    for k, v in vars(me).items():
        if not k.startswith("_") and k not in ("sys", "types", "torch", "install"):
            setattr(tf, k, v)
    linalg = types.ModuleType("tensorflow.linalg")
    linalg.logdet = _logdet
    linalg.inv = matrix_inverse
    linalg.cholesky = cholesky
    tf.linalg = linalg
    nn = types.ModuleType("tensorflow.nn")
    nn.relu = lambda x, name=None: torch.relu(_t(x))
    nn.sigmoid = sigmoid
    tf.nn = nn
    layers = types.ModuleType("tensorflow.layers")
    layers.flatten = lambda x: _t(x).reshape(x.shape[0], -1)
    tf.layers = layers
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.linalg"] = linalg

    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)

    import sklearn  # present in the image; only the removed `externals.joblib` is missing
    ext = types.ModuleType("sklearn.externals")
    import joblib as _joblib
    ext.joblib = _joblib
    sklearn.externals = ext
    sys.modules["sklearn.externals"] = ext
    sys.modules["sklearn.externals.joblib"] = _joblib
    return tf
