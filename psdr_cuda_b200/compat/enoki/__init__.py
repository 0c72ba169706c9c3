"""A small stand-in for the slice of Enoki's Python API that psdr-cuda's examples and docs use:
`enoki.cuda{,_autodiff}.{Float32, Vector3f, Matrix4f}`, arithmetic between them, `.numpy()`, and the module functions
`set_requires_gradient / forward / gradient / detach / slices / sqr / sqrt / hmean / squared_norm`.

It is NOT Enoki: values live in numpy, and differentiation is forward mode with respect to the variables marked by
`set_requires_gradient` (every array carries value + tangent). That is exactly what examples/run_test.py needs
(`ek.forward(P); ek.gradient(img)`): the tangent reaches the scene through `Mesh.set_transform` / `vertex_positions`,
`Integrator.renderD` returns an array whose tangent is filled by `pb_render_d_jvp` when `ek.forward` runs.
`ek.backward` is not provided — use the torch interface of `psdr_cuda` (`scene.parameter`, `Tensor.backward`) for reverse mode.
"""
import numpy as np

__psdr_b200_shim__ = True
_pending = []   # renderD outputs waiting for ek.forward


def _arr(x):
    return np.atleast_1d(np.asarray(x, dtype=np.float32))


def _tan(a, like):
    return np.zeros_like(like) if a is None else a


class Float32:
    def __init__(self, value=0.0, literal=False, tangent=None):
        if isinstance(value, Float32):
            self.v, self.d = value.v.copy(), None if value.d is None else value.d.copy()
        else:
            self.v, self.d = _arr(value), tangent
        self.requires_grad = False

    @staticmethod
    def zero(n=1):
        return Float32(np.zeros(n, np.float32))

    @staticmethod
    def full(value, n=1):
        return Float32(np.full(n, value, np.float32))

    def _coerce(self, o):
        return o if isinstance(o, Float32) else Float32(o)

    def __add__(self, o):
        o = self._coerce(o)
        d = None if self.d is None and o.d is None else _tan(self.d, self.v) + _tan(o.d, o.v)
        return Float32(self.v + o.v, tangent=d)
    __radd__ = __add__

    def __neg__(self):
        return Float32(-self.v, tangent=None if self.d is None else -self.d)

    def __sub__(self, o):
        return self + (-self._coerce(o))

    def __rsub__(self, o):
        return self._coerce(o) - self

    def __mul__(self, o):
        if isinstance(o, (Vector3f, Matrix4f)):
            return o.__rmul__(self)
        o = self._coerce(o)
        d = None if self.d is None and o.d is None else _tan(self.d, self.v) * o.v + self.v * _tan(o.d, o.v)
        return Float32(self.v * o.v, tangent=d)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._coerce(o)
        q = self.v / o.v
        d = None if self.d is None and o.d is None else (_tan(self.d, self.v) - q * _tan(o.d, o.v)) / o.v
        return Float32(q, tangent=d)

    def __len__(self):
        return len(self.v)

    def __getitem__(self, i):
        return float(self.v[i])

    def numpy(self):
        return self.v.copy()

    def __repr__(self):
        return "Float32(%s)" % np.array2string(self.v, threshold=8)


class Vector3f:
    def __init__(self, *args):
        if len(args) == 1 and isinstance(args[0], Vector3f):
            self.x, self.y, self.z = Float32(args[0].x), Float32(args[0].y), Float32(args[0].z)
        elif len(args) == 1:
            a = np.asarray(args[0], dtype=np.float32)
            if a.ndim == 1 and a.shape[0] == 3:
                self.x, self.y, self.z = Float32(a[0]), Float32(a[1]), Float32(a[2])
            else:
                a = a.reshape(-1, 3)
                self.x, self.y, self.z = Float32(a[:, 0].copy()), Float32(a[:, 1].copy()), Float32(a[:, 2].copy())
        elif len(args) == 3:
            self.x, self.y, self.z = (a if isinstance(a, Float32) else Float32(a) for a in args)
            self.x, self.y, self.z = Float32(self.x), Float32(self.y), Float32(self.z)
        elif len(args) == 0:
            self.x, self.y, self.z = Float32(0.0), Float32(0.0), Float32(0.0)
        else:
            raise TypeError("Vector3f: unsupported constructor arguments")

    @staticmethod
    def zero(n=1):
        return Vector3f(np.zeros((n, 3), np.float32))

    def _comps(self):
        return (self.x, self.y, self.z)

    def _zip(self, o, f):
        if isinstance(o, Vector3f):
            return Vector3f(*[f(a, b) for a, b in zip(self._comps(), o._comps())])
        return Vector3f(*[f(a, o) for a in self._comps()])

    def __add__(self, o):
        return self._zip(o, lambda a, b: a + b)
    __radd__ = __add__

    def __sub__(self, o):
        return self._zip(o, lambda a, b: a - b)

    def __mul__(self, o):
        return self._zip(o, lambda a, b: a * b)
    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._zip(o, lambda a, b: a / b)

    def __neg__(self):
        return Vector3f(-self.x, -self.y, -self.z)

    def __getitem__(self, i):
        return self._comps()[i]

    def __len__(self):
        return max(len(c) for c in self._comps())

    def numpy(self):
        n = len(self)
        return np.stack([np.broadcast_to(c.v, (n,)) for c in self._comps()], axis=1).astype(np.float32)

    def tangent_numpy(self):
        n = len(self)
        return np.stack([np.broadcast_to(_tan(c.d, c.v), (n,)) for c in self._comps()], axis=1).astype(np.float32)

    def has_tangent(self):
        return any(c.d is not None for c in self._comps())

    def __repr__(self):
        return "Vector3f(%s)" % np.array2string(self.numpy(), threshold=12)


class Matrix4f:
    """row-major 4x4 with tangent; `translate` / `rotate` follow enoki::translate / enoki::rotate (angle in radians, axis used as given)"""

    def __init__(self, value=None, tangent=None):
        self.v = np.eye(4, dtype=np.float32) if value is None else np.asarray(value, dtype=np.float32).reshape(4, 4)
        self.d = tangent

    @staticmethod
    def identity():
        return Matrix4f()

    @staticmethod
    def translate(vec):
        vec = vec if isinstance(vec, Vector3f) else Vector3f(vec)
        m = np.eye(4, dtype=np.float32)
        m[:3, 3] = vec.numpy()[0]
        d = None
        if vec.has_tangent():
            d = np.zeros((4, 4), np.float32)
            d[:3, 3] = vec.tangent_numpy()[0]
        return Matrix4f(m, d)

    @staticmethod
    def rotate(axis, angle):
        axis = axis if isinstance(axis, Vector3f) else Vector3f(axis)
        angle = angle if isinstance(angle, Float32) else Float32(angle)
        a = axis.numpy()[0].astype(np.float64)
        th = float(angle.v[0])
        s, c = np.sin(th), np.cos(th)

        def build(s, c, k, add_c):
            x, y, z = a
            m = np.array([[x * x * k + add_c, x * y * k - z * s, x * z * k + y * s, 0],
                          [y * x * k + z * s, y * y * k + add_c, y * z * k - x * s, 0],
                          [z * x * k - y * s, z * y * k + x * s, z * z * k + add_c, 0],
                          [0, 0, 0, 0]], dtype=np.float64)
            return m
        m = build(s, c, 1.0 - c, c)
        m[3, 3] = 1.0
        d = None
        if angle.d is not None:
            dth = float(angle.d[0])
            d = (build(c, -s, s, -s) * dth).astype(np.float32)   # d/dtheta of every entry
        return Matrix4f(m.astype(np.float32), d)

    def __matmul__(self, o):
        d = None
        if self.d is not None or o.d is not None:
            d = _tan(self.d, self.v) @ o.v + self.v @ _tan(o.d, o.v)
        return Matrix4f(self.v @ o.v, d)
    __mul__ = __matmul__

    def numpy(self):
        return self.v.copy()


# ---- module-level functions the examples use -----------------------------------------------------------------------------
def set_requires_gradient(x, flag=True):
    x.requires_grad = bool(flag)
    if isinstance(x, Float32):
        x.d = np.ones_like(x.v) if flag else None
    else:
        raise TypeError("this Enoki stand-in differentiates with respect to Float32 variables only")


def forward(x, free_graph=True):
    """push the tangent of `x` to everything computed from it; renderD outputs run their JVP now"""
    global _pending
    for out in _pending:
        out._run_forward()
    _pending = []


def gradient(y):
    if isinstance(y, Vector3f):
        return Vector3f(y.tangent_numpy())
    return Float32(_tan(y.d, y.v))


def detach(x):
    if isinstance(x, Vector3f):
        return Vector3f(x.numpy())
    if isinstance(x, Matrix4f):
        return Matrix4f(x.v.copy())
    return Float32(x.v.copy())


def backward(*a, **k):
    raise NotImplementedError("reverse mode goes through torch here: scene.parameter(...), img = integrator.renderD(...); loss.backward()")


def slices(x):
    return len(x)


def sqr(x):
    return x * x


def sqrt(x):
    x = x if isinstance(x, Float32) else Float32(x)
    r = np.sqrt(x.v)
    return Float32(r, tangent=None if x.d is None else x.d / (2 * r))


def hmean(x):
    if isinstance(x, Vector3f):
        return Float32(x.numpy().mean())
    return Float32(x.v.mean())


def squared_norm(v):
    return v.x * v.x + v.y * v.y + v.z * v.z
