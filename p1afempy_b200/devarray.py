"""Device-resident stand-in for the numpy arrays user callbacks receive.

The reference hands its coefficient callbacks (``f``, ``g``, ``a_ij``, ``phi``:
``BoundaryConditionType``, /root/reference/p1afempy/data_structures.py:43-48)
an ``(n, 2)`` array of points (or an ``(n,)`` array of values) and expects an
``(n,)`` array back.  At 64M triangles and a 16-point rule those arrays are
17 GB; shipping them over PCIe to evaluate ``sin`` on the host would cost
seconds, the assembly itself milliseconds.  ``DeviceArray`` wraps a CUDA tensor
and implements enough of the ndarray protocol (``__array_ufunc__``,
``__array_function__``, indexing, arithmetic, ``.shape``) that typical
vectorised callbacks -- e.g. ``np.sin(OMEGA*2.*r[:, 0])*np.sin(OMEGA*r[:, 1])``,
``-(1. + r[:, 0]**2)``, ``-np.ones(r.shape[0])`` -- run unchanged on the GPU.
Anything it cannot express (boolean-mask assignment, Python loops over entries,
third-party functions) raises ``DeviceArrayUnsupported``; the caller then
evaluates that callback on the host (only the callback: download points, call,
upload values).  Transcendentals evaluated on the device differ from numpy's
libm in the last bits; the assembly arithmetic itself stays bit-faithful.
"""
from __future__ import annotations

import numbers

import numpy as np
import torch


class DeviceArrayUnsupported(TypeError):
    pass


_UFUNCS = {
    np.add: torch.add, np.subtract: torch.subtract, np.multiply: torch.multiply,
    np.divide: torch.divide, np.true_divide: torch.divide, np.power: torch.pow,
    np.negative: torch.negative, np.positive: lambda x: x, np.absolute: torch.abs,
    np.fabs: torch.abs, np.sqrt: torch.sqrt, np.square: torch.square,
    np.exp: torch.exp, np.expm1: torch.expm1, np.log: torch.log, np.log1p: torch.log1p,
    np.log2: torch.log2, np.log10: torch.log10,
    np.sin: torch.sin, np.cos: torch.cos, np.tan: torch.tan,
    np.arcsin: torch.asin, np.arccos: torch.acos, np.arctan: torch.atan,
    np.arctan2: torch.atan2, np.hypot: torch.hypot,
    np.sinh: torch.sinh, np.cosh: torch.cosh, np.tanh: torch.tanh,
    np.maximum: torch.maximum, np.minimum: torch.minimum,
    np.floor: torch.floor, np.ceil: torch.ceil, np.sign: torch.sign,
    np.reciprocal: torch.reciprocal,
    np.less: torch.lt, np.less_equal: torch.le, np.greater: torch.gt,
    np.greater_equal: torch.ge, np.equal: torch.eq, np.not_equal: torch.ne,
    np.logical_and: torch.logical_and, np.logical_or: torch.logical_or,
    np.logical_not: torch.logical_not,
}


def _unwrap(x, like: torch.Tensor):
    if isinstance(x, DeviceArray):
        return x.t
    if isinstance(x, torch.Tensor):
        return x.to(like.device)
    if isinstance(x, numbers.Number):
        return x
    if isinstance(x, np.ndarray):
        if x.ndim == 0:
            return x.item()
        return torch.from_numpy(np.ascontiguousarray(x)).to(like.device)
    raise DeviceArrayUnsupported(type(x).__name__)


def _wrap(t):
    if isinstance(t, torch.Tensor):
        if t.dtype in (torch.float32, torch.float16, torch.bfloat16):
            t = t.to(torch.float64)
        return DeviceArray(t)
    return t


class DeviceArray:
    """Thin ndarray-like view of a CUDA tensor (float64 / bool / int)."""
    __array_priority__ = 1000

    def __init__(self, tensor: torch.Tensor):
        self.t = tensor

    # ---- shape protocol
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.ndim

    @property
    def size(self):
        return self.t.numel()

    @property
    def dtype(self):
        return torch.empty(0, dtype=self.t.dtype).numpy().dtype

    @property
    def T(self):
        return DeviceArray(self.t.T)

    def __len__(self):
        return self.t.shape[0]

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return DeviceArray(self.t.reshape(shape))

    def flatten(self):
        return DeviceArray(self.t.reshape(-1))

    def copy(self):
        return DeviceArray(self.t.clone())

    def astype(self, dtype):
        return DeviceArray(self.t.to(torch.from_numpy(np.empty(0, dtype=dtype)).dtype))

    # ---- indexing (basic only: r[:, 0], r[..., 1]); no item assignment
    def __getitem__(self, key):
        if isinstance(key, DeviceArray) or isinstance(key, np.ndarray):
            raise DeviceArrayUnsupported('fancy / boolean indexing')
        if isinstance(key, tuple) and any(isinstance(k, (DeviceArray, np.ndarray, list))
                                          for k in key):
            raise DeviceArrayUnsupported('fancy / boolean indexing')
        return _wrap(self.t[key])

    def __setitem__(self, key, value):
        raise DeviceArrayUnsupported('item assignment')

    def __iter__(self):
        raise DeviceArrayUnsupported('iteration over entries')

    def __array__(self, *args, **kwargs):
        # silent device->host conversion would defeat the purpose
        raise DeviceArrayUnsupported('implicit conversion to numpy')

    def __bool__(self):
        raise DeviceArrayUnsupported('truth value of an array')

    # ---- numpy protocols
    def __getattr__(self, name):
        # (only reached for attributes this class does not have)
        raise DeviceArrayUnsupported('ndarray attribute %r' % name)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__' or kwargs.get('out') is not None or ufunc not in _UFUNCS:
            raise DeviceArrayUnsupported('ufunc %s.%s' % (ufunc.__name__, method))
        args = [_unwrap(x, self.t) for x in inputs]
        if ufunc in (np.power,) and not isinstance(args[0], torch.Tensor):
            args[0] = torch.as_tensor(args[0], dtype=torch.float64, device=self.t.device)
        if ufunc in (np.maximum, np.minimum, np.arctan2, np.hypot):
            args = [a if isinstance(a, torch.Tensor)
                    else torch.as_tensor(a, dtype=torch.float64, device=self.t.device)
                    for a in args]
        try:
            return _wrap(_UFUNCS[ufunc](*args))
        except (RuntimeError, TypeError) as exc:   # what torch cannot broadcast / promote
            raise DeviceArrayUnsupported('%s: %s' % (ufunc.__name__, exc)) from None

    def __array_function__(self, func, types, args, kwargs):
        t = self.t
        if func in (np.zeros_like, np.ones_like, np.empty_like):
            fill = {np.zeros_like: 0., np.ones_like: 1., np.empty_like: 0.}[func]
            return DeviceArray(torch.full(t.shape, fill, dtype=torch.float64, device=t.device))
        if func is np.where and len(args) == 3:
            a, b, c = (_unwrap(x, t) for x in args)
            b = b if isinstance(b, torch.Tensor) else torch.as_tensor(
                b, dtype=torch.float64, device=t.device)
            c = c if isinstance(c, torch.Tensor) else torch.as_tensor(
                c, dtype=torch.float64, device=t.device)
            return _wrap(torch.where(a, b, c))
        if func is np.sum:
            axis = kwargs.get('axis', args[1] if len(args) > 1 else None)
            return _wrap(t.sum() if axis is None else t.sum(dim=axis))
        if func is np.shape:
            return self.shape
        raise DeviceArrayUnsupported('numpy function %s' % func.__name__)

    # ---- arithmetic
    def _bin(self, other, op, swap=False):
        o = _unwrap(other, self.t)
        a, b = (o, self.t) if swap else (self.t, o)
        if not isinstance(a, torch.Tensor):
            a = torch.as_tensor(a, dtype=torch.float64, device=self.t.device)
        try:
            return _wrap(op(a, b))
        except (RuntimeError, TypeError) as exc:   # what torch cannot broadcast / promote
            raise DeviceArrayUnsupported('%s: %s' % (getattr(op, '__name__', 'op'), exc)) from None

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.subtract)
    def __rsub__(self, o): return self._bin(o, torch.subtract, True)
    def __mul__(self, o): return self._bin(o, torch.multiply)
    def __rmul__(self, o): return self._bin(o, torch.multiply, True)
    def __truediv__(self, o): return self._bin(o, torch.divide)
    def __rtruediv__(self, o): return self._bin(o, torch.divide, True)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __rpow__(self, o): return self._bin(o, torch.pow, True)
    def __neg__(self): return _wrap(torch.negative(self.t))
    def __pos__(self): return self
    def __abs__(self): return _wrap(torch.abs(self.t))
    def __lt__(self, o): return self._bin(o, torch.lt)
    def __le__(self, o): return self._bin(o, torch.le)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __ge__(self, o): return self._bin(o, torch.ge)
    def __eq__(self, o): return self._bin(o, torch.eq)      # noqa: PLW1641
    def __ne__(self, o): return self._bin(o, torch.ne)
    def __and__(self, o): return self._bin(o, torch.logical_and)
    def __or__(self, o): return self._bin(o, torch.logical_or)
    def __invert__(self): return _wrap(torch.logical_not(self.t))
    __hash__ = None


def evaluate(fn, points: torch.Tensor, n: int):
    """Calls a user callback on device data.

    points: CUDA tensor (n, 2) or (n,).  Returns a float64 CUDA tensor (n,), or
    None when the callback cannot run on the proxy (the caller evaluates it on
    the host instead).  A callback that builds its result on the host from
    ``.shape[0]`` (``-np.ones(n)``) is fine: the values are uploaded."""
    try:
        out = fn(DeviceArray(points))
    except DeviceArrayUnsupported:
        return None     # (any other exception is the callback's own: it propagates)
    dev = points.device
    if isinstance(out, DeviceArray):
        t = out.t.to(torch.float64)
    elif isinstance(out, torch.Tensor):
        t = out.to(device=dev, dtype=torch.float64)
    elif isinstance(out, numbers.Number):
        t = torch.full((n,), float(out), dtype=torch.float64, device=dev)
    else:
        try:
            host = np.asarray(out, dtype=np.float64)
        except (TypeError, ValueError):
            return None
        if host.ndim == 0:
            t = torch.full((n,), float(host), dtype=torch.float64, device=dev)
        else:
            t = torch.from_numpy(np.ascontiguousarray(host)).to(dev)
    t = t.reshape(-1)
    if t.numel() == 1 and n != 1:
        t = t.expand(n)
    if t.numel() != n:
        raise ValueError('callback returned %d values for %d points' % (t.numel(), n))
    return t.contiguous()
