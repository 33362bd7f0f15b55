"""Eager executor for the TensorFlow symbols the reference imports, on torch-CPU - TEST INFRASTRUCTURE (oracle/tfemu/README.md).

Put ``oracle/tfemu`` on ``sys.path`` BEFORE importing anything of ``/root/reference/sac_eo``: ``import tensorflow as tf``
then resolves here and the reference's unmodified classes run.  Only what the reference uses exists; anything else
raises AttributeError.  Semantics restated from TensorFlow 2.x (the reference pins no version; these have been stable
across 2.x):

* binary operators between a Tensor and a NumPy array / Python scalar convert the non-tensor operand to the TENSOR's
  dtype (``ops.convert_to_tensor(y, dtype_hint=x.dtype)``), so ``float32 tensor * float64 ndarray`` is a float32
  product; two tensors of different float dtypes raise (TF: InvalidArgumentError).  ``__array_priority__ = 100`` makes
  ``ndarray <op> Tensor`` dispatch to the tensor's reflected operator, as for ``EagerTensor``.
* ``tf.concat`` / ``tf.reduce_min`` over a Python sequence convert every element to the dtype of the first Tensor in
  it (``args_to_matching_eager``).
* ``GradientTape.gradient`` = reverse-mode autodiff of everything executed since the tape was opened; a source that is
  not connected yields ``None``.  Nested tapes differentiate through the inner gradient (trpo.py:213-222).
* gradients at ties / boundaries: ``reduce_min`` / ``reduce_max`` split the cotangent equally between equal entries
  (``_MinOrMaxGrad``); ``maximum(x, y)`` passes it to ``x`` where ``x >= y`` (``_MaximumMinimumGrad``);
  ``clip_by_value`` passes it where ``lo <= x <= hi`` (closed interval, ``_ClipByValueGrad``); ``abs`` uses sign(x);
  relu passes it where ``x > 0``.
* ``tf.keras.optimizers.Adam`` (β1 0.9, β2 0.999, ε 1e-7), dense update of ``keras/optimizers/adam.py::update_step``
  (TF >= 2.11; the reference's shipped log TEMPLOG_0 is dated 2023-05):  ``α = lr·sqrt(1-β2^t)/(1-β1^t)`` in the
  variable's dtype, ``m += (g-m)·(1-β1)``, ``v += (g²-v)·(1-β2)``, ``θ -= m·α / (sqrt(v)+ε)`` - ε is NOT bias-corrected,
  and ``1-β`` is formed in Python double and THEN rounded to the variable's dtype (fp32(0.001) = 0.00100000005).  The
  fused ``ApplyAdam`` kernel of TF <= 2.10 forms ``T(1) - beta2`` in fp32 instead (0.00099998713): v differs by 1.3e-5
  relative, a step by 6e-6 - ``Adam(legacy_one_minus_beta=True)`` selects that variant (tests bound the difference).
* ``Dense``: ``act(x @ kernel + bias)``, kernel ``[in, units]``, bias zeros; ``Sequential.get_weights()`` returns
  ``[k0, b0, k1, b1, …]`` as NumPy copies.
"""
import builtins
import math as _math
import sys
import types

import numpy as _np
import torch as _torch

__version__ = "2.emu"
_torch.set_grad_enabled(True)


# ---------------------------------------------------------------------------------------------------------------------
# dtypes
# ---------------------------------------------------------------------------------------------------------------------
class DType:
    def __init__(self, name, tdtype, npdtype):
        self.name, self._t, self.as_numpy_dtype = name, tdtype, npdtype

    @property
    def base_dtype(self):
        return self

    @property
    def is_floating(self):
        return self._t.is_floating_point

    def __repr__(self):
        return "tf." + self.name

    def __eq__(self, other):
        return isinstance(other, DType) and other._t == self._t or (isinstance(other, str) and other == self.name)

    def __hash__(self):
        return hash(self.name)


float16 = DType("float16", _torch.float16, _np.float16)
float32 = DType("float32", _torch.float32, _np.float32)
float64 = DType("float64", _torch.float64, _np.float64)
int32 = DType("int32", _torch.int32, _np.int32)
int64 = DType("int64", _torch.int64, _np.int64)
bool = DType("bool", _torch.bool, _np.bool_)          # noqa: A001  (tf.bool)
_BY_TORCH = {d._t: d for d in (float16, float32, float64, int32, int64, bool)}


def _tdtype(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, DType):
        return dtype._t
    if isinstance(dtype, _torch.dtype):
        return dtype
    return _torch.from_numpy(_np.zeros(0, _np.dtype(dtype))).dtype


# ---------------------------------------------------------------------------------------------------------------------
# tensors
# ---------------------------------------------------------------------------------------------------------------------
class TensorShape(tuple):
    def as_list(self):
        return list(self)

    @property
    def rank(self):
        return len(self)


def _raw(x, hint=None):
    """-> torch tensor.  Tensors keep their dtype; NumPy / Python values take ``hint`` when it is floating and the value
    is real-valued (convert_to_tensor with dtype_hint), else their own NumPy dtype (float -> float32 like tf.constant)."""
    if isinstance(x, Tensor):
        return x._t
    if isinstance(x, _torch.Tensor):
        return x
    if isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in _flatten(x)):
        first = next(e for e in _flatten(x) if isinstance(e, Tensor))
        return _torch.stack([_raw(e, first._t.dtype) for e in x])
    a = _np.asarray(x)
    if a.dtype == object:
        raise TypeError("cannot convert %r to a tensor" % (x,))
    t = _torch.from_numpy(_np.ascontiguousarray(a)) if a.ndim else _torch.tensor(a.item())
    if hint is not None:
        if hint.is_floating_point and (t.dtype.is_floating_point or t.dtype in (_torch.int32, _torch.int64, _torch.bool)):
            # Python ints / floats and NumPy floats follow the tensor operand; NumPy INT arrays would make TF raise,
            # the reference never relies on that
            return t.to(hint)
        if not hint.is_floating_point and not t.dtype.is_floating_point:
            return t.to(hint)
        return t
    if isinstance(x, (float, builtins.int)) or (isinstance(x, (list, tuple)) and a.dtype == _np.float64):
        return t.to(_torch.float32) if a.dtype == _np.float64 else t.to(_torch.int32)      # tf.constant defaults
    return t


def _flatten(x):
    for e in x:
        if isinstance(e, (list, tuple)):
            yield from _flatten(e)
        else:
            yield e


def _pair(a, b):
    """Operands of a binary op as raw tensors (TF's dtype rule, see the module docstring)."""
    if isinstance(a, Tensor) and isinstance(b, Tensor):
        if a._t.dtype != b._t.dtype:
            raise TypeError("InvalidArgumentError: cannot compute binary op: dtypes %s and %s differ"
                            % (a._t.dtype, b._t.dtype))
        return a._t, b._t
    if isinstance(a, Tensor):
        return a._t, _raw(b, a._t.dtype)
    return _raw(a, b._t.dtype), b._t


class Tensor:
    """EagerTensor stand-in around a torch tensor that may carry an autograd history."""
    __array_priority__ = 100

    def __init__(self, t):
        self._t = t

    # -- introspection ------------------------------------------------------------------------------------------------
    @property
    def shape(self):
        return TensorShape(self._t.shape)

    @property
    def dtype(self):
        return _BY_TORCH[self._t.dtype]

    @property
    def ndim(self):
        return self._t.dim()

    def numpy(self):
        a = self._t.detach().cpu().numpy().copy()
        return a if a.ndim else a[()]

    def __array__(self, dtype=None, copy=None):
        a = self._t.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a.copy()

    def __len__(self):
        if self._t.dim() == 0:
            raise TypeError("Scalar tensor has no len()")
        return self._t.shape[0]

    def __iter__(self):
        if self._t.dim() == 0:
            raise TypeError("Cannot iterate over a scalar tensor")
        return (Tensor(self._t[i]) for i in range(self._t.shape[0]))

    def __float__(self):
        return float(self._t.detach())

    def __int__(self):
        return builtins.int(self._t.detach())

    def __bool__(self):
        return builtins.bool(self._t.detach())

    def __repr__(self):
        return "<tfemu.Tensor shape=%s dtype=%s numpy=%r>" % (tuple(self._t.shape), self.dtype.name, self.numpy())

    def __hash__(self):
        return id(self)

    def __getitem__(self, idx):
        if isinstance(idx, Tensor):
            idx = idx._t
        elif isinstance(idx, tuple):
            idx = tuple(i._t if isinstance(i, Tensor) else i for i in idx)
        return Tensor(self._t[idx])

    # -- arithmetic ---------------------------------------------------------------------------------------------------
    def __add__(self, o):
        a, b = _pair(self, o); return Tensor(a + b)

    def __radd__(self, o):
        a, b = _pair(o, self); return Tensor(a + b)

    def __sub__(self, o):
        a, b = _pair(self, o); return Tensor(a - b)

    def __rsub__(self, o):
        a, b = _pair(o, self); return Tensor(a - b)

    def __mul__(self, o):
        a, b = _pair(self, o); return Tensor(a * b)

    def __rmul__(self, o):
        a, b = _pair(o, self); return Tensor(a * b)

    def __truediv__(self, o):
        a, b = _pair(self, o); return Tensor(a / b)

    def __rtruediv__(self, o):
        a, b = _pair(o, self); return Tensor(a / b)

    def __pow__(self, o):
        a, b = _pair(self, o); return Tensor(a ** b)

    def __rpow__(self, o):
        a, b = _pair(o, self); return Tensor(a ** b)

    def __matmul__(self, o):
        a, b = _pair(self, o); return Tensor(a @ b)

    def __neg__(self):
        return Tensor(-self._t)

    def __abs__(self):
        return abs(self)

    def __lt__(self, o):
        a, b = _pair(self, o); return Tensor(a < b)

    def __le__(self, o):
        a, b = _pair(self, o); return Tensor(a <= b)

    def __gt__(self, o):
        a, b = _pair(self, o); return Tensor(a > b)

    def __ge__(self, o):
        a, b = _pair(self, o); return Tensor(a >= b)

    def __eq__(self, o):
        try:
            a, b = _pair(self, o)
        except TypeError:
            return NotImplemented
        return Tensor(a == b)

    def __ne__(self, o):
        a, b = _pair(self, o); return Tensor(a != b)


class Variable(Tensor):
    """``tf.Variable``: a leaf the tapes differentiate with respect to; ``assign*`` write in place without history."""

    def __init__(self, initial_value, dtype=None, name=None, trainable=True):
        t = _raw(initial_value, None)
        td = _tdtype(dtype)
        if td is not None:
            t = t.to(td)
        super().__init__(t.detach().clone().requires_grad_(t.dtype.is_floating_point))
        self.name = name or "Variable"
        self.trainable = trainable

    def _write(self, value, fn):
        v = _raw(value, self._t.dtype)
        if v.dtype != self._t.dtype:
            v = v.to(self._t.dtype)
        with _torch.no_grad():
            fn(self._t, v.detach().reshape(self._t.shape) if v.numel() == self._t.numel() else v.detach())
        return self

    def assign(self, value):
        return self._write(value, lambda t, v: t.copy_(v))

    def assign_add(self, value):
        return self._write(value, lambda t, v: t.add_(v))

    def assign_sub(self, value):
        return self._write(value, lambda t, v: t.sub_(v))

    def value(self):
        return Tensor(self._t)

    def read_value(self):
        return Tensor(self._t)


def _T(x, hint=None):
    return x if isinstance(x, Tensor) else Tensor(_raw(x, hint))


def _arg(x):
    """Argument of a unary op / reduction as a raw torch tensor (convert_to_tensor WITHOUT a dtype hint): tensors as they
    are, Python scalars as float32, NumPy arrays and NumPy scalars in their own dtype (tf.square(np.float64 array) stays
    float64), sequences that contain Tensors stacked in the first Tensor's dtype."""
    if isinstance(x, Tensor):
        return x._t
    a = _np.asarray(x) if not isinstance(x, (list, tuple)) or not any(isinstance(e, Tensor) for e in _flatten(x)) else None
    if a is None:
        return _raw(x)
    if isinstance(x, (float, builtins.int)):
        return _torch.tensor(x, dtype=_torch.float32)
    return _torch.from_numpy(_np.ascontiguousarray(a)) if a.ndim else _torch.tensor(a.item(), dtype=_tdtype(a.dtype))


# ---------------------------------------------------------------------------------------------------------------------
# tapes
# ---------------------------------------------------------------------------------------------------------------------
_TAPES = []


class GradientTape:
    def __init__(self, persistent=False, watch_accessed_variables=True):
        self.persistent = persistent

    def __enter__(self):
        _TAPES.append(self)
        return self

    def __exit__(self, *exc):
        _TAPES.remove(self)
        return False

    def watch(self, x):
        for e in (x if isinstance(x, (list, tuple)) else [x]):
            if not e._t.requires_grad:
                e._t.requires_grad_(True)

    def gradient(self, target, sources, output_gradients=None):
        """Sources may be a variable, a list, or a nested list (trpo.py:67 passes ``[actor.trainable, alpha]``); the
        result has the same structure, ``None`` where a source is not connected to the target."""
        srcs = list(_flatten(sources)) if isinstance(sources, (list, tuple)) else [sources]
        outer = any(t is not self for t in _TAPES)          # an enclosing tape must see the gradient computation
        tgt = target._t
        og = None
        if output_gradients is not None:
            og = _raw(output_gradients, tgt.dtype)
        elif tgt.dim() > 0:
            og = _torch.ones_like(tgt)
        if not tgt.requires_grad:
            grads = [None] * len(srcs)
        else:
            grads = _torch.autograd.grad(tgt, [s._t for s in srcs], grad_outputs=og, retain_graph=True,
                                         create_graph=outer, allow_unused=True)
        it = iter(None if g is None else Tensor(g if outer else g.detach()) for g in grads)

        def pack(struct):
            if isinstance(struct, (list, tuple)):
                return [pack(e) for e in struct]
            return next(it)
        return pack(sources)


def stop_gradient(x):
    return Tensor(_arg(x).detach())


def function(fn=None, **kw):
    return fn if fn is not None else (lambda f: f)


# ---------------------------------------------------------------------------------------------------------------------
# ops
# ---------------------------------------------------------------------------------------------------------------------
def _axis(axis):
    return None if axis is None else (tuple(axis) if isinstance(axis, (list, tuple)) else builtins.int(axis))


def _reduce(fn, x, axis, keepdims):
    t = _arg(x)
    ax = _axis(axis)
    if ax is None:
        r = fn(t)
        return Tensor(r.reshape([1] * t.dim()) if keepdims else r)
    return Tensor(fn(t, dim=ax, keepdim=keepdims))


def reduce_sum(x, axis=None, keepdims=False):
    return _reduce(_torch.sum, x, axis, keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    return _reduce(_torch.mean, x, axis, keepdims)


def reduce_min(x, axis=None, keepdims=False):
    return _reduce(_torch.amin, x, axis, keepdims)           # amin / amax: equal split between ties, like TF


def reduce_max(x, axis=None, keepdims=False):
    return _reduce(_torch.amax, x, axis, keepdims)


def reduce_logsumexp(x, axis=None, keepdims=False):
    t = _arg(x)
    ax = _axis(axis)
    return Tensor(_torch.logsumexp(t, dim=tuple(range(t.dim())) if ax is None else ax, keepdim=keepdims))


def reduce_euclidean_norm(x, axis=None, keepdims=False):
    t = _arg(x)
    return Tensor(_torch.sqrt(_torch.sum(t * t, dim=_axis(axis), keepdim=keepdims)) if axis is not None
                  else _torch.sqrt(_torch.sum(t * t)))


def norm(x, ord="euclidean", axis=None, keepdims=False):
    assert ord in ("euclidean", 2)
    return reduce_euclidean_norm(x, axis, keepdims)


def square(x):
    t = _arg(x); return Tensor(t * t)


def exp(x):
    return Tensor(_torch.exp(_arg(x)))


def log(x):
    return Tensor(_torch.log(_arg(x)))


def sqrt(x):
    return Tensor(_torch.sqrt(_arg(x)))


def abs(x):                                                   # noqa: A001
    return Tensor(_torch.abs(_arg(x)))


def tanh(x):
    return Tensor(_torch.tanh(_arg(x)))


def atanh(x):
    return Tensor(_torch.atanh(_arg(x)))


def softplus(x):
    return Tensor(_softplus(_arg(x)))


def _softplus(t):
    # log(1 + exp(t)) evaluated stably the way Eigen's softplus functor does: max(t, 0) + log1p(exp(-|t|))
    return _torch.clamp(t, min=0) + _torch.log1p(_torch.exp(-_torch.abs(t)))


class _MaxGrad(_torch.autograd.Function):
    """maximum(x, y): cotangent to x where x >= y, else to y (TF ``_MaximumMinimumGrad``; torch would halve ties)."""

    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x >= y)
        ctx.sx, ctx.sy = x.shape, y.shape
        return _torch.maximum(x, y)

    @staticmethod
    def backward(ctx, g):
        (m,) = ctx.saved_tensors
        z = _torch.zeros_like(g)
        return _unbroadcast(_torch.where(m, g, z), ctx.sx), _unbroadcast(_torch.where(m, z, g), ctx.sy)


def _unbroadcast(g, shape):
    while g.dim() > len(shape):
        g = g.sum(0)
    for i, n in enumerate(shape):
        if n == 1 and g.shape[i] != 1:
            g = g.sum(i, keepdim=True)
    return g


def maximum(x, y):
    if isinstance(x, Tensor) or isinstance(y, Tensor):
        a, b = _pair(x, y)
    else:
        a, b = _arg(x), _arg(y)
        b = b.to(a.dtype)
    return Tensor(_MaxGrad.apply(a, b))


def minimum(x, y):
    return -maximum(-_T(x), -_T(y))


class _ClipGrad(_torch.autograd.Function):
    """clip_by_value: cotangent passes where lo <= x <= hi (closed interval, TF ``_ClipByValueGrad``)."""

    @staticmethod
    def forward(ctx, x, lo, hi):
        ctx.save_for_backward((x >= lo) & (x <= hi))
        return _torch.minimum(_torch.maximum(x, lo), hi)

    @staticmethod
    def backward(ctx, g):
        (m,) = ctx.saved_tensors
        return _torch.where(m, g, _torch.zeros_like(g)), None, None


def clip_by_value(t, clip_value_min, clip_value_max):
    x = _arg(t)
    lo = _raw(clip_value_min, x.dtype).detach()
    hi = _raw(clip_value_max, x.dtype).detach()
    return Tensor(_ClipGrad.apply(x, lo, hi))


def global_norm(t_list):
    ts = [_arg(t) for t in t_list if t is not None]
    return Tensor(_torch.sqrt(sum(_torch.sum(t * t) for t in ts)))


def clip_by_global_norm(t_list, clip_norm, use_norm=None):
    """``clip_ops.clip_by_global_norm``: t * clip_norm / max(global_norm, clip_norm)."""
    gn = global_norm(t_list) if use_norm is None else _T(use_norm)
    c = _raw(clip_norm, gn._t.dtype)
    scale = c * _torch.minimum(1.0 / gn._t, 1.0 / c)
    return [None if t is None else Tensor(_arg(t) * scale) for t in t_list], gn


def squeeze(x, axis=None):
    t = _arg(x)
    if axis is None:
        return Tensor(t.squeeze())
    ax = _axis(axis)
    for a in ([ax] if isinstance(ax, builtins.int) else ax):
        if t.shape[a] != 1:
            raise ValueError("InvalidArgumentError: can not squeeze dim[%d], expected a dimension of 1, got %d"
                             % (a, t.shape[a]))
    return Tensor(t.squeeze(ax))


def expand_dims(x, axis):
    return Tensor(_arg(x).unsqueeze(builtins.int(axis)))


def reshape(x, shape):
    return Tensor(_arg(x).reshape([builtins.int(s) for s in shape]))


def _seq(values):
    vals = list(values)
    first = next((v for v in vals if isinstance(v, Tensor)), None)
    if first is None:
        ts = [_arg(v) for v in vals]
        return [t.to(ts[0].dtype) for t in ts]
    return [_raw(v, first._t.dtype) if not isinstance(v, Tensor) else v._t for v in vals]


def concat(values, axis):
    ts = _seq(values)
    if any(t.dtype != ts[0].dtype for t in ts):
        raise TypeError("InvalidArgumentError: concat of mixed dtypes %s" % [t.dtype for t in ts])
    return Tensor(_torch.cat(ts, dim=builtins.int(axis)))


def stack(values, axis=0):
    return Tensor(_torch.stack(_seq(values), dim=builtins.int(axis)))


def split(value, num_or_size_splits, axis=0):
    t = _arg(value)
    if isinstance(num_or_size_splits, builtins.int):
        assert t.shape[axis] % num_or_size_splits == 0
        return [Tensor(p) for p in _torch.split(t, t.shape[axis] // num_or_size_splits, dim=axis)]
    return [Tensor(p) for p in _torch.split(t, list(num_or_size_splits), dim=axis)]


def cast(x, dtype):
    td = _tdtype(dtype)
    if isinstance(x, Tensor):
        return x if x._t.dtype == td else Tensor(x._t.to(td))
    a = _np.asarray(x)
    t = _torch.from_numpy(_np.ascontiguousarray(a)) if a.ndim else _torch.tensor(a.item())
    return Tensor(t.to(td))


def constant(value, dtype=None, shape=None):
    t = _raw(value)
    if dtype is not None:
        t = t.to(_tdtype(dtype))
    return Tensor(t.reshape(shape) if shape is not None else t)


def convert_to_tensor(value, dtype=None, dtype_hint=None):
    if isinstance(value, Tensor):
        return value
    return constant(value, dtype)


def _shape_arg(shape):
    if isinstance(shape, (builtins.int, _np.integer)):
        return [builtins.int(shape)]
    return [builtins.int(s) for s in shape]


def ones(shape, dtype=float32):
    return Tensor(_torch.ones(_shape_arg(shape), dtype=_tdtype(dtype)))


def zeros(shape, dtype=float32):
    return Tensor(_torch.zeros(_shape_arg(shape), dtype=_tdtype(dtype)))


def ones_like(x, dtype=None):
    return Tensor(_torch.ones_like(_arg(x).detach(), dtype=_tdtype(dtype)))


def zeros_like(x, dtype=None):
    return Tensor(_torch.zeros_like(_arg(x).detach(), dtype=_tdtype(dtype)))


def shape(x):
    return Tensor(_torch.tensor(list(_arg(x).shape), dtype=_torch.int32))


def size(x):
    return Tensor(_torch.tensor(_arg(x).numel(), dtype=_torch.int32))


def argmax(x, axis=None):
    return Tensor(_torch.argmax(_arg(x), dim=0 if axis is None else builtins.int(axis)))


def one_hot(indices, depth, dtype=float32):
    i = _raw(indices).to(_torch.int64)
    return Tensor(_torch.nn.functional.one_hot(i, builtins.int(depth)).to(_tdtype(dtype)))


def matmul(a, b):
    x, y = _pair(_T(a), b) if not isinstance(b, Tensor) or isinstance(a, Tensor) else _pair(a, b)
    return Tensor(x @ y)


# ---------------------------------------------------------------------------------------------------------------------
# tf.math / tf.nn / tf.linalg / tf.random / tf.config
# ---------------------------------------------------------------------------------------------------------------------
def _module(name, **members):
    m = types.ModuleType(__name__ + "." + name)
    m.__dict__.update(members)
    sys.modules[m.__name__] = m
    return m


def _relu(x):
    return Tensor(_torch.relu(_arg(x)))


def _elu(x):
    return Tensor(_torch.nn.functional.elu(_arg(x)))


def _softmax_xent(labels, logits, axis=-1):
    lg = _arg(logits)
    return Tensor(-_torch.sum(_raw(labels, lg.dtype) * _torch.log_softmax(lg, dim=axis), dim=axis))


math = _module("math", log=log, exp=exp, sqrt=sqrt, square=square, softplus=softplus, tanh=tanh, atanh=atanh, abs=abs,
               reduce_euclidean_norm=reduce_euclidean_norm, reduce_sum=reduce_sum, reduce_mean=reduce_mean,
               reduce_min=reduce_min, reduce_max=reduce_max, reduce_logsumexp=reduce_logsumexp, maximum=maximum,
               minimum=minimum, argmax=argmax)
nn = _module("nn", softplus=softplus, tanh=tanh, relu=_relu, elu=_elu,
             softmax_cross_entropy_with_logits=_softmax_xent)
linalg = _module("linalg", global_norm=global_norm, norm=norm, matmul=matmul)

_SEED = [0]


def _set_seed(seed):
    _SEED[0] = builtins.int(seed)
    _torch.manual_seed(builtins.int(seed))


random = _module("random", set_seed=_set_seed)
_experimental = _module("config.experimental", list_physical_devices=lambda kind=None: [],
                        set_memory_growth=lambda dev, enable: None)
config = _module("config", experimental=_experimental, list_physical_devices=lambda kind=None: [])


# ---------------------------------------------------------------------------------------------------------------------
# tf.keras
# ---------------------------------------------------------------------------------------------------------------------
class _Initializer:
    def _rng(self):
        _SEED[0] += 1
        return _np.random.default_rng(_SEED[0])


class Orthogonal(_Initializer):
    """QR of a Gaussian, sign-fixed, times ``gain`` (keras/initializers: Orthogonal)."""

    def __init__(self, gain=1.0, seed=None):
        self.gain = gain

    def __call__(self, shape, dtype=None):
        rows, cols = shape
        a = self._rng().standard_normal((builtins.max(rows, cols), builtins.min(rows, cols)))
        q, r = _np.linalg.qr(a)
        q = q * _np.sign(_np.diag(r))
        if rows < cols:
            q = q.T
        return (self.gain * q[:rows, :cols]).astype(_np.float32)


class VarianceScaling(_Initializer):
    def __init__(self, scale=1.0, mode="fan_in", distribution="truncated_normal", seed=None):
        self.scale, self.mode, self.distribution = scale, mode, distribution

    def __call__(self, shape, dtype=None):
        fan_in, fan_out = shape
        n = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2.0}[self.mode]
        s = self.scale / builtins.max(1.0, n)
        if self.distribution == "uniform":
            lim = _math.sqrt(3.0 * s)
            return self._rng().uniform(-lim, lim, shape).astype(_np.float32)
        return (self._rng().standard_normal(shape) * _math.sqrt(s)).astype(_np.float32)


def _get_initializer(spec):
    if isinstance(spec, str):
        if spec == "glorot_uniform":
            return VarianceScaling(1.0, "fan_avg", "uniform")
        raise ValueError("initializer %r not emulated" % spec)
    return spec


class _Layer:
    built = False

    @property
    def trainable_variables(self):
        return []

    def build(self, in_dim):
        self.built = True
        return in_dim


class Dense(_Layer):
    def __init__(self, units, kernel_initializer="glorot_uniform", activation=None, input_shape=None, use_bias=True,
                 name=None):
        self.units, self.activation = builtins.int(units), activation
        self._init = _get_initializer(kernel_initializer)
        self.input_shape_arg = input_shape
        self.kernel = self.bias = None

    def build(self, in_dim):
        self.kernel = Variable(self._init((builtins.int(in_dim), self.units)), dtype=float32, name="kernel")
        self.bias = Variable(_np.zeros(self.units, _np.float32), dtype=float32, name="bias")
        self.built = True
        return self.units

    @property
    def trainable_variables(self):
        return [self.kernel, self.bias]

    def __call__(self, x):
        y = Tensor(_arg(x) @ self.kernel._t + self.bias._t)
        return self.activation(y) if self.activation is not None else y


class LayerNormalization(_Layer):
    def __init__(self, axis=-1, epsilon=1e-3):
        self.eps = epsilon
        self.gamma = self.beta = None

    def build(self, in_dim):
        self.gamma = Variable(_np.ones(in_dim, _np.float32), dtype=float32, name="gamma")
        self.beta = Variable(_np.zeros(in_dim, _np.float32), dtype=float32, name="beta")
        self.built = True
        return in_dim

    @property
    def trainable_variables(self):
        return [self.gamma, self.beta]

    def __call__(self, x):
        t = _arg(x)
        mu = t.mean(-1, keepdim=True)
        var = ((t - mu) ** 2).mean(-1, keepdim=True)
        return Tensor((t - mu) * _torch.rsqrt(var + self.eps) * self.gamma._t + self.beta._t)


class Activation(_Layer):
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, x):
        return self.fn(x)


class Sequential:
    """Layers are built as they are added when the first one names its ``input_shape`` (the reference always does,
    nn_utils.py:60-66), so ``trainable_variables`` is complete right after ``create_nn``."""

    def __init__(self, layers=None, name=None):
        self.name = name
        self.layers = []
        self._dim = None
        for l in layers or []:
            self.add(l)

    def add(self, layer):
        if not self.layers and getattr(layer, "input_shape_arg", None) is not None:
            self._dim = builtins.int(layer.input_shape_arg[0])
        self.layers.append(layer)
        if self._dim is not None and not layer.built:
            self._dim = layer.build(self._dim)

    def _ensure_built(self, x):
        if self._dim is None:
            self._dim = builtins.int(x.shape[-1])
            for l in self.layers:
                if not l.built:
                    self._dim = l.build(self._dim)

    @property
    def trainable_variables(self):
        return [v for l in self.layers for v in l.trainable_variables]

    trainable_weights = trainable_variables
    variables = trainable_variables
    weights = trainable_variables

    def __call__(self, x, training=None):
        x = _T(x)
        self._ensure_built(x)
        for l in self.layers:
            x = l(x)
        return x

    def get_weights(self):
        return [v.numpy() for v in self.trainable_variables]

    def set_weights(self, weights):
        vs = self.trainable_variables
        if len(weights) != len(vs):
            raise ValueError("set_weights: %d arrays for %d variables" % (len(weights), len(vs)))
        for v, w in zip(vs, weights):
            w = _np.asarray(w)
            if tuple(w.shape) != tuple(v.shape):
                raise ValueError("set_weights: shape %s for variable %s" % (w.shape, tuple(v.shape)))
            v.assign(w)


class Adam:
    """Dense Keras Adam (module docstring).  ``slots[id(var)] = (m, v)``; ``iterations`` counts apply_gradients calls.
    ``last_grads`` keeps what the last call received (the golden generator records it)."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, name="Adam",
                 legacy_one_minus_beta=False, **kw):
        assert not amsgrad
        self.legacy_one_minus_beta = legacy_one_minus_beta
        # a float32 variable like Keras' (ppo.py:109-118 reads it with .numpy() and re-assigns it)
        self.learning_rate = Variable(_np.float32(learning_rate), dtype=float32, name="learning_rate", trainable=False)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.iterations = 0
        self.slots = {}
        self.last_grads = None

    lr = property(lambda self: self.learning_rate)

    def apply_gradients(self, grads_and_vars):
        gv = [(g, v) for g, v in grads_and_vars]
        self.last_grads = [None if g is None else _np.array(_arg(g).detach().numpy()) for g, _ in gv]
        self.iterations += 1
        t = self.iterations
        for g, var in gv:
            if g is None:
                continue
            dt = var._t.dtype
            g = _arg(g).detach().to(dt).reshape(var._t.shape)
            if id(var) not in self.slots:
                self.slots[id(var)] = (_torch.zeros_like(var._t), _torch.zeros_like(var._t))
            m, v = self.slots[id(var)]
            lr = self.learning_rate._t.detach().to(dt)
            b1 = _torch.tensor(self.beta_1, dtype=dt)
            b2 = _torch.tensor(self.beta_2, dtype=dt)
            one = _torch.tensor(1.0, dtype=dt)
            b1p = _torch.pow(b1, _torch.tensor(float(t), dtype=dt))
            b2p = _torch.pow(b2, _torch.tensor(float(t), dtype=dt))
            alpha = lr * _torch.sqrt(one - b2p) / (one - b1p)
            if self.legacy_one_minus_beta:
                omb1, omb2 = one - b1, one - b2
            else:
                omb1 = _torch.tensor(1 - self.beta_1, dtype=dt)
                omb2 = _torch.tensor(1 - self.beta_2, dtype=dt)
            with _torch.no_grad():
                m.add_((g - m) * omb1)
                v.add_((g * g - v) * omb2)
                var._t.sub_((m * alpha) / (_torch.sqrt(v) + _torch.tensor(self.epsilon, dtype=dt)))

    def get_slot_arrays(self, var):
        m, v = self.slots[id(var)]
        return m.numpy().copy(), v.numpy().copy()


_activations = _module("keras.activations", tanh=tanh, relu=_relu, elu=_elu)
_initializers = _module("keras.initializers", Orthogonal=Orthogonal, VarianceScaling=VarianceScaling)
_layers = _module("keras.layers", Dense=Dense, LayerNormalization=LayerNormalization, Activation=Activation)
_optimizers = _module("keras.optimizers", Adam=Adam)
keras = _module("keras", activations=_activations, initializers=_initializers, layers=_layers, optimizers=_optimizers,
                Sequential=Sequential)
