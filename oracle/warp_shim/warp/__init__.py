"""TEST INFRASTRUCTURE ONLY -- a minimal pure-Python stand-in for NVIDIA Warp (`warp-lang`, pyproject.toml:37 of the reference,
third-party, not installable here), just large enough to EXECUTE the reference's own `@wp.kernel` / `@wp.func` Python bodies on
the CPU, one thread at a time, so that golden vectors for the Warp-implemented parts of the hot path (scene collision, tool pose,
c-space costs) come from the reference's source instead of from a restatement (tests/golden/make_warp_golden.py).

What is the reference's: every kernel and function body (imported from /root/reference, never copied).  What is restated here:
Warp's scalar typing rules (int / int truncates, float32 arithmetic) and ~35 builtins with the semantics of Warp's documentation
(quaternions are xyzw; transform = (p, q); transform_point(t, x) = quat_rotate(q, x) + p; sign(0) = +1; quat_inverse = conjugate).
Nothing here is fast and nothing outside tests/ imports it.
"""
import inspect
import math
import sys
import threading
import types as _pytypes

import numpy as np

_tls = threading.local()


# ------------------------------------------------------------------------------------------------ scalars
class _Int(int):
    """C-like integer: int / int truncates toward zero (Warp kernels are statically typed)."""

    def _w(self, v):
        return type(self)(v)

    def __add__(self, o): return self._w(int(self) + int(o)) if isinstance(o, int) else NotImplemented
    def __radd__(self, o): return self._w(int(o) + int(self)) if isinstance(o, int) else NotImplemented
    def __sub__(self, o): return self._w(int(self) - int(o)) if isinstance(o, int) else NotImplemented
    def __rsub__(self, o): return self._w(int(o) - int(self)) if isinstance(o, int) else NotImplemented
    def __mul__(self, o): return self._w(int(self) * int(o)) if isinstance(o, int) else NotImplemented
    def __rmul__(self, o): return self._w(int(o) * int(self)) if isinstance(o, int) else NotImplemented
    def __neg__(self): return self._w(-int(self))

    @staticmethod
    def _cdiv(a, b):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q

    def __truediv__(self, o): return self._w(self._cdiv(int(self), int(o))) if isinstance(o, int) else NotImplemented
    def __rtruediv__(self, o): return self._w(self._cdiv(int(o), int(self))) if isinstance(o, int) else NotImplemented
    __floordiv__ = __truediv__
    __rfloordiv__ = __rtruediv__

    def __mod__(self, o):
        if not isinstance(o, int):
            return NotImplemented
        a, b = int(self), int(o)
        return self._w(a - self._cdiv(a, b) * b)

    def __and__(self, o): return self._w(int(self) & int(o))
    def __or__(self, o): return self._w(int(self) | int(o))
    def __xor__(self, o): return self._w(int(self) ^ int(o))
    def __lshift__(self, o): return self._w(int(self) << int(o))
    def __rshift__(self, o): return self._w(int(self) >> int(o))


def _int_type(name):
    def __new__(cls, v=0):
        if isinstance(v, (float, np.floating)):
            v = math.trunc(float(v))
        return int.__new__(cls, int(v))
    return type(name, (_Int,), {"__new__": __new__})


int8, int16, int32, int64 = (_int_type(n) for n in ("int8", "int16", "int32", "int64"))
uint8, uint16, uint32, uint64 = (_int_type(n) for n in ("uint8", "uint16", "uint32", "uint64"))
float16, float32, float64 = np.float16, np.float32, np.float64
bool = bool  # noqa: A001


def _f(x):
    return np.float32(x)


# ------------------------------------------------------------------------------------------------ vectors
class _Vec:
    n = 0
    dt = np.float32
    __array_ufunc__ = None   # numpy scalars must defer to __rmul__ etc. instead of broadcasting over the object

    def __init__(self, *a):
        if len(a) == 0:
            self.v = np.zeros(self.n, self.dt)
        elif len(a) == 1 and isinstance(a[0], _Vec):
            self.v = a[0].v.astype(self.dt).copy()
        elif len(a) == 1 and isinstance(a[0], (np.ndarray, list, tuple)):
            self.v = np.asarray(a[0], self.dt).reshape(self.n).copy()
        elif len(a) == 1:
            self.v = np.full(self.n, a[0], self.dt)
        else:
            flat = []
            for x in a:                                   # vec4(vec3, w) style constructors
                flat.extend(list(x.v) if isinstance(x, _Vec) else [x])
            assert len(flat) == self.n, (type(self).__name__, a)
            self.v = np.asarray(flat, self.dt)

    def __getitem__(self, i): return self.v[int(i)] if self.dt is np.float32 else int32(self.v[int(i)])
    def __setitem__(self, i, x): self.v[int(i)] = x
    def _new(self, arr):
        o = type(self).__new__(type(self))
        o.v = np.asarray(arr, self.dt)
        return o
    def __add__(self, o): return self._new(self.v + o.v)
    def __sub__(self, o): return self._new(self.v - o.v)
    def __neg__(self): return self._new(-self.v)
    def __mul__(self, s): return self._new(self.v * self.dt(s)) if not isinstance(s, _Vec) else NotImplemented
    def __rmul__(self, s): return self._new(self.dt(s) * self.v)
    def __truediv__(self, s): return self._new(self.v / self.dt(s))
    def __eq__(self, o): return isinstance(o, _Vec) and np.array_equal(self.v, o.v)
    def __len__(self): return self.n
    def __repr__(self): return f"{type(self).__name__}({', '.join(str(x) for x in self.v)})"
    x = property(lambda s: s.v[0], lambda s, val: s.v.__setitem__(0, val))
    y = property(lambda s: s.v[1], lambda s, val: s.v.__setitem__(1, val))
    z = property(lambda s: s.v[2], lambda s, val: s.v.__setitem__(2, val))
    w = property(lambda s: s.v[3], lambda s, val: s.v.__setitem__(3, val))


def _vec_type(name, n, dt=np.float32, base=_Vec):
    return type(name, (base,), {"n": n, "dt": dt})


vec2, vec3, vec4 = _vec_type("vec2", 2), _vec_type("vec3", 3), _vec_type("vec4", 4)
vec2f, vec3f, vec4f = vec2, vec3, vec4
vec2i, vec3i, vec4i = (_vec_type(f"vec{k}i", k, np.int32) for k in (2, 3, 4))
_vector_cache = {}


def vector(*values, length=None, dtype=float32):
    """wp.vector(length=n, dtype=t) is a type; wp.vector(a, b, c, ...) builds a value."""
    if length is None:
        return vector(length=len(values), dtype=float32)(*values)
    key = (int(length), dtype)
    if key not in _vector_cache:
        std = {(2, np.float32): vec2, (3, np.float32): vec3, (4, np.float32): vec4}.get(key)
        _vector_cache[key] = std or _vec_type(f"vec{length}", int(length), np.float32 if dtype is np.float32 else np.int32)
    return _vector_cache[key]


class quat(_Vec):  # xyzw
    n, dt = 4, np.float32

    def __mul__(self, o):
        if isinstance(o, quat):
            a, b = self.v, o.v
            return quat(a[3] * b[0] + b[3] * a[0] + a[1] * b[2] - b[1] * a[2],
                        a[3] * b[1] + b[3] * a[1] + a[2] * b[0] - b[2] * a[0],
                        a[3] * b[2] + b[3] * a[2] + a[0] * b[1] - b[0] * a[1],
                        a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2])
        return _Vec.__mul__(self, o)


quatf = quat


def quaternion(*values, dtype=float32):
    """wp.quaternion(dtype=t) is a type; wp.quaternion(x, y, z, w) builds a value."""
    return quat(*values) if values else quat



class transform:
    def __init__(self, *a):
        if len(a) == 0:
            self.p, self.q = vec3(), quat(0.0, 0.0, 0.0, 1.0)
        elif len(a) == 2:
            self.p, self.q = vec3(a[0]), quat(a[1])
        elif len(a) == 7:
            self.p, self.q = vec3(*a[:3]), quat(*a[3:])
        else:
            raise TypeError(a)

    def __getitem__(self, i):
        i = int(i)
        return self.p[i] if i < 3 else self.q[i - 3]

    def __mul__(self, o): return transform_multiply(self, o)
    def __repr__(self): return f"transform({self.p}, {self.q})"


transformf = transform
transformation = lambda dtype=float32: transform  # noqa: E731


def dot(a, b): return _f(np.dot(a.v, b.v))
def cross(a, b): return vec3(np.cross(a.v, b.v).astype(np.float32))
def length(a): return _f(np.sqrt(_f(np.dot(a.v, a.v))))
def length_sq(a): return _f(np.dot(a.v, a.v))
def normalize(a):
    n = length(a)
    return a._new(a.v / n) if n > 0 else a._new(np.zeros_like(a.v))
def cw_mul(a, b): return a._new(a.v * b.v)
def cw_div(a, b): return a._new(a.v / b.v)


def mul(a, b):
    return a * b


def quat_inverse(q): return quat(-q.v[0], -q.v[1], -q.v[2], q.v[3])
def quat_identity(): return quat(0.0, 0.0, 0.0, 1.0)


def quat_rotate(q, x):
    # warp/native/quat.h: x (2 w^2 - 1) + 2 w (q.xyz cross x) + 2 q.xyz (q.xyz . x)
    qv = vec3(q.v[:3])
    w = q.v[3]
    c = _f(2.0) * w * w - _f(1.0)
    d = _f(2.0) * dot(qv, x)
    return x * c + cross(qv, x) * (w * _f(2.0)) + qv * d


def quat_rotate_inv(q, x):
    return quat_rotate(quat_inverse(q), x)


def transform_point(t, p): return quat_rotate(t.q, p) + t.p
def transform_vector(t, v): return quat_rotate(t.q, v)
def transform_get_translation(t): return vec3(t.p)
def transform_get_rotation(t): return quat(t.q)
def transform_identity(): return transform()


def transform_inverse(t):
    qi = quat_inverse(t.q)
    return transform(-quat_rotate(qi, t.p), qi)


def transform_multiply(a, b):
    return transform(quat_rotate(a.q, b.p) + a.p, a.q * b.q)


# ------------------------------------------------------------------------------------------------ scalar builtins
def _num(x):
    return x


def abs(x): return type(x)(-x) if x < 0 else x  # noqa: A001
def min(a, b): return a if a < b else b  # noqa: A001
def max(a, b): return a if a > b else b  # noqa: A001
def clamp(x, lo, hi): return min(max(x, lo), hi)
def sign(x): return type(x)(-1) if x < 0 else type(x)(1)
def sqrt(x): return _f(np.sqrt(_f(x)))
def sin(x): return _f(np.sin(_f(x)))
def cos(x): return _f(np.cos(_f(x)))
def tan(x): return _f(np.tan(_f(x)))
def acos(x): return _f(np.arccos(_f(x)))
def asin(x): return _f(np.arcsin(_f(x)))
def atan2(y, x): return _f(np.arctan2(_f(y), _f(x)))
def exp(x): return _f(np.exp(_f(x)))
def log(x): return _f(np.log(_f(x)))
def floor(x): return _f(np.floor(_f(x)))
def ceil(x): return _f(np.ceil(_f(x)))
def round(x): return _f(np.sign(_f(x)) * np.floor(np.abs(_f(x)) + _f(0.5)))  # noqa: A001  (C roundf: half away from zero)
def pow(x, y): return _f(np.power(_f(x), _f(y)))  # noqa: A001
def isnan(x): return np.isnan(x)
def isinf(x): return np.isinf(x)
def select(cond, a, b): return b if cond else a
def where(cond, a, b): return a if cond else b


# ------------------------------------------------------------------------------------------------ arrays
class array:
    """wp.array(dtype=...) in an annotation -> a placeholder; array(data=np.ndarray, dtype=...) -> storage."""

    def __init__(self, data=None, dtype=None, ndim=1, **kw):
        self.dtype = dtype
        self.ndim = ndim
        self.data = None
        if data is not None:
            a = np.ascontiguousarray(data)
            if isinstance(dtype, type) and issubclass(dtype, _Vec):
                a = a.reshape(-1, dtype.n) if ndim == 1 else a.reshape(a.shape[0], -1, dtype.n)
            self.data = a
            self.shape = a.shape[:ndim] if not (isinstance(dtype, type) and issubclass(dtype, _Vec)) else a.shape[:-1]

    def _wrap(self, x):
        dt = self.dtype
        if isinstance(dt, type) and issubclass(dt, _Vec):
            return dt(np.array(x))
        if isinstance(dt, type) and issubclass(dt, _Int):
            return dt(int(x))
        return x  # numpy scalar of the array's own dtype

    def _idx(self, i):
        return tuple(int(k) for k in i) if isinstance(i, tuple) else int(i)

    def __getitem__(self, i):
        i = self._idx(i)
        n = self.data.shape[0] if not isinstance(i, tuple) else None
        if n is not None and not (0 <= i < n):
            raise IndexError(f"warp shim: index {i} out of range {n}")
        return self._wrap(self.data[i])

    def __setitem__(self, i, v):
        i = self._idx(i)
        self.data[i] = v.v if isinstance(v, _Vec) else v

    def numpy(self): return self.data


def array1d(dtype=None, **k): return array(dtype=dtype, ndim=1)
def array2d(dtype=None, **k): return array(dtype=dtype, ndim=2)
def array3d(dtype=None, **k): return array(dtype=dtype, ndim=3)
def array4d(dtype=None, **k): return array(dtype=dtype, ndim=4)


def from_numpy(a, dtype=None, ndim=None, **k):
    a = np.ascontiguousarray(a)
    if ndim is None:
        ndim = a.ndim - (1 if isinstance(dtype, type) and issubclass(dtype, _Vec) and a.ndim > 1 else 0)
        ndim = builtins_max(ndim, 1)
    return array(a, dtype=dtype, ndim=ndim)


def from_torch(t, dtype=None, **k):
    return from_numpy(t.detach().cpu().numpy(), dtype=dtype)


import builtins as _b  # noqa: E402
builtins_max = _b.max


def atomic_add(arr, i, v=None, *rest):
    if rest:                                             # 2-D form atomic_add(arr, i, j, v)
        i, v = (i, v), rest[0]
    old = arr[i]
    arr[i] = (old + v) if not isinstance(old, _Vec) else old + v
    return old


def atomic_min(arr, i, v):
    old = arr[i]
    arr[i] = min(old, v)
    return old


def atomic_max(arr, i, v):
    old = arr[i]
    arr[i] = max(old, v)
    return old


# ------------------------------------------------------------------------------------------------ functions, kernels, structs
class Function:
    """@wp.func: callable; several definitions under one name in one module are overloads, picked by the annotated type of
    the first parameter (the reference registers per-obstacle-type overloads this way, geom/collision/wp_collision_kernel.py:45-60)."""

    def __init__(self, name):
        self.name = name
        self.overloads = []

    def add(self, fn):
        self.overloads.append(fn)
        self.func = fn
        return self

    def __call__(self, *a, **k):
        if len(self.overloads) == 1:
            return self.overloads[0](*a, **k)
        for fn in self.overloads:
            params = list(inspect.signature(fn).parameters.values())
            ann = params[0].annotation if params else None
            name = ann if isinstance(ann, str) else getattr(ann, "__name__", None)   # `from __future__ import annotations`
            if name is not None and name.split(".")[-1] == type(a[0]).__name__:
                return fn(*a, **k)
        raise TypeError(f"warp shim: no overload of {self.name} for {type(a[0]).__name__}")


_functions = {}


def func(f=None, *, module=None, **kw):
    if f is None:
        return lambda g: func(g, module=module, **kw)
    if isinstance(f, Function):
        f = f.func
    key = (module or f.__module__, f.__name__)
    fn = _functions.get(key) or Function(f.__name__)
    _functions[key] = fn
    return fn.add(f)


class Kernel:
    def __init__(self, f=None, func=None, key=None, module=None, **kw):
        f = f if f is not None else func
        self.func = f
        self.key = f.__name__
        self.__name__ = f.__name__

    def __call__(self, *a, **k):
        return self.func(*a, **k)


def kernel(f=None, **kw):
    if f is None:
        return lambda g: Kernel(g)
    return Kernel(f)


def struct(cls):
    ann = dict(getattr(cls, "__annotations__", {}))

    def __init__(self, **kw):
        for k in ann:
            setattr(self, k, kw.get(k))
    cls.__init__ = __init__
    return cls


def constant(x): return x
def static(x): return x


def tid():
    t = _tls.tid
    return t if isinstance(t, tuple) else int32(t)


def _coerce(val, ann):
    if isinstance(ann, type) and issubclass(ann, _Int) and isinstance(val, (int, np.integer)):
        return ann(int(val))
    if ann is np.float32 and isinstance(val, (int, float, np.floating)) and not isinstance(val, _b.bool):
        return np.float32(val)
    return val


def launch(kernel, dim, inputs=(), outputs=(), device=None, stream=None, **kw):  # noqa: A002
    fn = kernel.func if isinstance(kernel, Kernel) else kernel
    args = list(inputs) + list(outputs)
    params = list(inspect.signature(fn).parameters.values())
    assert len(args) == len(params), f"{fn.__name__}: {len(args)} arguments for {len(params)} parameters"
    args = [_coerce(a, p.annotation) for a, p in zip(args, params)]
    dims = (dim,) if isinstance(dim, (int, np.integer)) else tuple(dim)
    total = int(np.prod(dims))
    for t in range(total):
        if len(dims) == 1:
            _tls.tid = t
        else:
            _tls.tid = tuple(int32(i) for i in np.unravel_index(t, dims))
        fn(*args)


# ------------------------------------------------------------------------------------------------ everything else: inert
class _Inert:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Inert()
    def __getattr__(self, n): return _Inert()
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def __iter__(self): return iter(())
    def __bool__(self): return False


def __getattr__(name):  # wp.config, wp.init, wp.ScopedTimer, wp.context, wp.Mesh, ... : host-side plumbing the goldens never run
    if name.startswith("__"):
        raise AttributeError(name)
    return _Inert()


def _submodule(name):
    m = _pytypes.ModuleType(name)
    m.__getattr__ = lambda n: _Inert()
    sys.modules[name] = m
    return m


for _n in ("warp.types", "warp.context", "warp.torch", "warp.config", "warp.utils", "warp.sim", "warp.build"):
    _submodule(_n)
