"""A minimal, serial, pure-Python/numpy stand-in for the `taichi` module.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py).  Taichi is not installable in
the build image, so the reference's *own, unmodified* Python sources (core/partice_system/*.py,
core/sph/*.py of jiajun-c/Ti-SPH) are executed on top of this module to produce golden vectors.
Only the part of the Taichi API those files touch is provided.

Semantics that matter for the golden vectors:
  * `float`/`ti.f32` fields are numpy float32, `int`/`ti.i32` fields are int32; every field read
    yields an f32/i32 scalar, Python scalars are "weak" (NEP 50), so kernel arithmetic runs in
    IEEE binary32 -- Taichi's default_fp/default_ip -- without fused multiply-add and without
    fast-math;
  * `@ti.kernel` / `@ti.func` bodies run as plain Python: every parallel `for` is executed
    serially in index order, so atomics have their serial meaning (stable counting sort);
  * scalar field elements are handed out as references (`Ref`), so `ti.template()` arguments
    such as `ret` in `for_all_neighbors(p_i, task, self.ps.density[p_i])` and the operand of
    `ti.atomic_add/sub` are l-values as in Taichi; plain locals passed to un-annotated
    arguments stay by-value (Taichi >= 1.1), which is what makes SURVEY quirk Q5 visible;
  * vector field elements are handed out as copies that write through on component
    assignment (`self.ps.x[p_i][0] = ...`).
"""
import numpy as np

f32 = np.float32
f64 = np.float64
i32 = np.int32
int32 = np.int32
i64 = np.int64
cpu = "cpu"
cuda = "cuda"
gpu = "gpu"
i, j, k, l = "i", "j", "k", "l"
ij, ijk = "ij", "ijk"


def _dtype(dt):
    if dt is float or dt is f32:
        return np.float32
    if dt is int or dt is i32:
        return np.int32
    return np.dtype(dt).type


def init(*args, **kwargs):
    return None


def data_oriented(cls):
    return cls


def kernel(fn):
    return fn


def func(fn):
    return fn


def static(x):
    return x


def template():
    return None


class types:
    @staticmethod
    def ndarray(*a, **k):
        return None


def _val(x):
    return x.v if isinstance(x, Ref) else x


class Ref:
    """l-value of one scalar field element."""
    __array_ufunc__ = None
    __slots__ = ("a", "i")

    def __init__(self, a, i):
        self.a, self.i = a, i

    @property
    def v(self):
        return self.a[self.i]

    # conversions
    def __index__(self): return int(self.v)
    def __int__(self): return int(self.v)
    def __float__(self): return float(self.v)
    def __bool__(self): return bool(self.v)
    def __hash__(self): return hash(self.v)
    def __repr__(self): return repr(self.v)
    # arithmetic (value semantics)
    def __add__(self, o): return self.v + _val(o)
    def __radd__(self, o): return _val(o) + self.v
    def __sub__(self, o): return self.v - _val(o)
    def __rsub__(self, o): return _val(o) - self.v
    def __mul__(self, o): return self.v * _val(o)
    def __rmul__(self, o): return _val(o) * self.v
    def __truediv__(self, o): return self.v / _val(o)
    def __rtruediv__(self, o): return _val(o) / self.v
    def __pow__(self, o): return self.v ** _val(o)
    def __neg__(self): return -self.v
    def __eq__(self, o): return self.v == _val(o)
    def __ne__(self, o): return self.v != _val(o)
    def __lt__(self, o): return self.v < _val(o)
    def __le__(self, o): return self.v <= _val(o)
    def __gt__(self, o): return self.v > _val(o)
    def __ge__(self, o): return self.v >= _val(o)
    # in-place (reference semantics)
    def __iadd__(self, o):
        self.a[self.i] = self.v + _val(o)
        return self

    def __isub__(self, o):
        self.a[self.i] = self.v - _val(o)
        return self

    def __imul__(self, o):
        self.a[self.i] = self.v * _val(o)
        return self

    def __itruediv__(self, o):
        self.a[self.i] = self.v / _val(o)
        return self


class Vector(np.ndarray):
    """ti.Vector value: small f32/i32 numpy array."""

    def __new__(cls, data, dt=None):
        a = np.asarray([_val(d) for d in data] if not isinstance(data, np.ndarray) else data)
        if dt is not None:
            a = a.astype(_dtype(dt))
        elif a.dtype == np.float64:
            a = a.astype(np.float32)
        elif a.dtype == np.int64:
            a = a.astype(np.int32)
        return a.view(cls)

    def __array_finalize__(self, obj):
        self._home = None

    @staticmethod
    def zero(dt, n):
        return Vector(np.zeros(n, _dtype(dt)))

    @staticmethod
    def field(n, dtype=float, shape=None):
        return Field(_dtype(dtype), shape, n)

    def norm(self):
        s = self[0] * self[0]
        for q in range(1, len(self)):
            s = s + self[q] * self[q]
        return np.sqrt(s)

    def dot(self, o):
        s = self[0] * o[0]
        for q in range(1, len(self)):
            s = s + self[q] * o[q]
        return s

    def cast(self, dt):
        return Vector(np.asarray(self).astype(_dtype(dt)))

    def __getitem__(self, idx):
        r = np.ndarray.__getitem__(self, idx)
        return r[()] if isinstance(r, np.ndarray) and r.ndim == 0 else r

    def __setitem__(self, idx, value):
        np.ndarray.__setitem__(self, idx, _val(value))
        home = getattr(self, "_home", None)
        if home is not None:                      # write-through: field[p][c] = value
            home[0][home[1]][idx] = _val(value)


def _key(idx):
    if isinstance(idx, tuple):
        out = []
        for q in idx:
            q = _val(q)
            if isinstance(q, np.ndarray):
                out.extend(int(t) for t in q)
            else:
                out.append(int(q))
        return tuple(out)
    idx = _val(idx)
    if idx is None:
        return ()
    if isinstance(idx, np.ndarray):
        return tuple(int(t) for t in idx)
    return int(idx)


class Field:
    def __init__(self, dtype, shape=None, vec=0):
        self.dtype, self.vec = dtype, vec
        self.arr = None
        self.shape = None
        if shape is not None:
            self._alloc(shape)

    def _alloc(self, shape):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        self.shape = tuple(int(s) for s in shape)
        full = self.shape + ((self.vec,) if self.vec else ())
        self.arr = np.zeros(full, self.dtype)

    def __getitem__(self, idx):
        kk = _key(idx)
        if self.vec:
            v = Vector(self.arr[kk].copy())
            v._home = (self.arr, kk)
            return v
        return Ref(self.arr, kk)

    def __setitem__(self, idx, value):
        kk = _key(idx)
        value = _val(value)
        self.arr[kk] = value

    def fill(self, value):
        self.arr.fill(value)

    def to_numpy(self):
        return self.arr.copy()

    def from_numpy(self, a):
        self.arr[...] = a


def field(dtype, shape=None):
    return Field(_dtype(dtype), shape)


class _Node:
    def __init__(self, shape):
        self.shape = tuple(shape)

    def dense(self, axes, shape):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        return _Node(self.shape + tuple(int(s) for s in shape))

    def place(self, *fields):
        for f in fields:
            f._alloc(self.shape)


root = _Node(())


def grouped(x):
    if isinstance(x, Field):
        if len(x.shape) == 1:
            return range(x.shape[0])
        return (Vector(np.array(t, np.int32)) for t in np.ndindex(*x.shape))
    return x


def ndrange(*ranges):
    spans = [range(*r) if isinstance(r, tuple) else range(r) for r in ranges]

    def gen():
        import itertools
        for t in itertools.product(*spans):          # row-major: first index slowest
            yield Vector(np.array(t, np.int32))
    return gen()


def atomic_add(ref, v):
    old = ref.v
    ref.a[ref.i] = old + _val(v)
    return old


def atomic_sub(ref, v):
    old = ref.v
    ref.a[ref.i] = old - _val(v)
    return old


def max(a, b):     # noqa: A001
    return np.maximum(_val(a), _val(b))


def min(a, b):     # noqa: A001
    return np.minimum(_val(a), _val(b))


def pow(a, b):     # noqa: A001
    # scalar `**` goes through numpy's scalar-math path = libm powf for float32; np.power (the
    # ufunc) may dispatch to a vectorised SVML pow with different last-bit rounding
    return _val(a) ** _val(b)


def sqrt(a):
    return np.sqrt(_val(a))


class algorithms:
    class PrefixSumExecutor:
        """inclusive, in-place i32 scan (Taichi's ti.algorithms.PrefixSumExecutor.run)."""

        def __init__(self, n):
            self.n = n

        def run(self, f):
            f.arr[...] = np.cumsum(f.arr, dtype=np.int64).astype(np.int32)
