"""A small stand-in for the slice of Enoki's Python API that psdr-cuda's examples and docs use:
`enoki.cuda{,_autodiff}.{Float32, Vector3f, Matrix4f}`, arithmetic between them, `.numpy()`, and the module functions
`set_requires_gradient / forward / backward / gradient / detach / slices / sqr / sqrt / hsum / hmean / squared_norm / ...`.

It is NOT Enoki: values live in numpy on the host; only the renderer runs on the GPU. Both AD modes of the reference's scripts work:

* forward mode (examples/run_test.py:122-129, `P = FloatD(0.); ek.set_requires_gradient(P); ...; ek.forward(P); ek.gradient(img)`):
  every array carries value + tangent (dual numbers); the tangent reaches the scene through `Mesh.set_transform` /
  `vertex_positions` / texture data, and `Integrator.renderD` returns an array whose tangent is filled by `pb_render_d_jvp` when
  `ek.forward` runs. A forward seed is attached to size-1 `Float32` variables (the scripts' scalar parameter `P`).
* reverse mode (docs/inverse_diff_render.rst:63-79, examples/utils/adam.py:26-65, `ek.set_requires_gradient(mesh.vertex_positions);
  img = integrator.renderD(scene); loss = ...; ek.backward(loss); ek.gradient(mesh.vertex_positions)`): operations on arrays that
  depend on a variable marked by `set_requires_gradient` are recorded on a tape; `backward` walks it in reverse, and where it meets
  the output of `renderD` it hands dL/dI to `pb_render_d_vjp` (the hand-written reverse-mode kernels) and distributes the flat
  gradient vector to the arrays that were stored in the scene.
"""
import numpy as np

__psdr_b200_shim__ = True
_pending = []   # renderD outputs waiting for ek.forward
_counter = [0]  # creation order of taped arrays (= a topological order)


def _arr(x):
    return np.atleast_1d(np.asarray(x, dtype=np.float32))


def _tan(a, like):
    return np.zeros_like(like) if a is None else a


def _unbroadcast(g, shape):
    """adjoint of numpy broadcasting: sum `g` back to `shape` (arrays here are 1-D, length 1 or n)"""
    g = np.asarray(g, dtype=np.float64)
    if g.shape == tuple(shape):
        return g
    if tuple(shape) == (1,):
        return np.atleast_1d(g.sum())
    return np.broadcast_to(g, shape).copy()


class Float32:
    """1-D float array with an optional forward tangent `d` and an optional tape entry `_bw = (parents, fn)`; `fn(adj)` returns one
    adjoint per parent"""

    def __init__(self, value=0.0, literal=False, tangent=None):
        self._bw = None
        self._grad = None
        self._adj = None
        self.requires_grad = False
        _counter[0] += 1
        self._id = _counter[0]
        if isinstance(value, Float32):
            self.v, self.d = value.v.copy(), None if value.d is None else value.d.copy()
            if value._tracked():   # a copy stays connected to its source (Enoki copies share the variable)
                self._bw = ((value,), lambda g: (g,))
        else:
            self.v, self.d = _arr(value), tangent

    def _tracked(self):
        return self.requires_grad or self._bw is not None

    @staticmethod
    def _make(v, d, parents, fn):
        out = Float32(v, tangent=d)
        parents = tuple(parents)
        if any(p._tracked() for p in parents):
            out._bw = (parents, fn)
        return out

    @staticmethod
    def zero(n=1):
        return Float32(np.zeros(n, np.float32))

    @staticmethod
    def full(value, n=1):
        return Float32(np.full(n, value, np.float32))

    def _coerce(self, o):
        return o if isinstance(o, Float32) else Float32(o)

    def __add__(self, o):
        if isinstance(o, (Vector3f, Matrix4f)):
            return o.__radd__(self)
        o = self._coerce(o)
        d = None if self.d is None and o.d is None else _tan(self.d, self.v) + _tan(o.d, o.v)
        return Float32._make(self.v + o.v, d, (self, o), lambda g: (g, g))
    __radd__ = __add__

    def __neg__(self):
        return Float32._make(-self.v, None if self.d is None else -self.d, (self,), lambda g: (-g,))

    def __sub__(self, o):
        if isinstance(o, Vector3f):
            return (-o) + self
        o = self._coerce(o)
        d = None if self.d is None and o.d is None else _tan(self.d, self.v) - _tan(o.d, o.v)
        return Float32._make(self.v - o.v, d, (self, o), lambda g: (g, -g))

    def __rsub__(self, o):
        return self._coerce(o) - self

    def __mul__(self, o):
        if isinstance(o, (Vector3f, Matrix4f)):
            return o.__rmul__(self)
        o = self._coerce(o)
        a, b = self.v, o.v
        d = None if self.d is None and o.d is None else _tan(self.d, a) * b + a * _tan(o.d, b)
        return Float32._make(a * b, d, (self, o), lambda g: (g * b, g * a))
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._coerce(o)
        a, b = self.v, o.v
        q = a / b
        d = None if self.d is None and o.d is None else (_tan(self.d, a) - q * _tan(o.d, b)) / b
        return Float32._make(q, d, (self, o), lambda g: (g / b, -g * q / b))

    def __rtruediv__(self, o):
        return self._coerce(o) / self

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
            self.x, self.y, self.z = (Float32(a) for a in args)
        elif len(args) == 0:
            self.x, self.y, self.z = Float32(0.0), Float32(0.0), Float32(0.0)
        else:
            raise TypeError("Vector3f: unsupported constructor arguments")

    @staticmethod
    def _from_comps(x, y, z):
        v = Vector3f.__new__(Vector3f)
        v.x, v.y, v.z = x, y, z
        return v

    @staticmethod
    def zero(n=1):
        return Vector3f(np.zeros((n, 3), np.float32))

    @property
    def requires_grad(self):
        return any(c.requires_grad for c in self._comps())

    def _tracked(self):
        return any(c._tracked() for c in self._comps())

    def _comps(self):
        return (self.x, self.y, self.z)

    def _zip(self, o, f):
        if isinstance(o, Vector3f):
            return Vector3f._from_comps(*[f(a, b) for a, b in zip(self._comps(), o._comps())])
        return Vector3f._from_comps(*[f(a, o) for a in self._comps()])

    def __add__(self, o):
        return self._zip(o, lambda a, b: a + b)
    __radd__ = __add__

    def __sub__(self, o):
        return self._zip(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._zip(o, lambda a, b: b - a)

    def __mul__(self, o):
        return self._zip(o, lambda a, b: a * b)
    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._zip(o, lambda a, b: a / b)

    def __neg__(self):
        return Vector3f._from_comps(-self.x, -self.y, -self.z)

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
    """row-major 4x4 with tangent; `translate` / `rotate` follow enoki::translate / enoki::rotate (angle in radians, axis used as given).
    Forward mode only (the scripts differentiate transforms with ek.forward)."""

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
def _leaves(x):
    if isinstance(x, Vector3f):
        return list(x._comps())
    if isinstance(x, Float32):
        return [x]
    raise TypeError("this Enoki stand-in differentiates Float32 and Vector3f arrays")


def set_requires_gradient(x, flag=True):
    """mark `x` as a differentiable variable. Reverse mode: any array. Forward mode additionally seeds the tangent of a size-1 Float32
    (the scalar parameter of examples/run_test.py) with one."""
    for c in _leaves(x):
        c.requires_grad = bool(flag)
        c._grad = None
        if isinstance(x, Float32) and len(c.v) == 1:
            c.d = np.ones_like(c.v) if flag else None


def requires_gradient(x):
    return any(c.requires_grad for c in _leaves(x))


def forward(x, free_graph=True):
    """push the tangent of `x` to everything computed from it; renderD outputs run their JVP now"""
    global _pending
    for out in _pending:
        out._run_forward()
    _pending = []


def backward(loss, free_graph=True):
    """reverse-mode sweep from `loss` (a Float32; several entries are summed): fills what `gradient` returns for every array marked
    by `set_requires_gradient` the loss depends on. Outputs of `Integrator.renderD` on the way are differentiated by the renderer's
    reverse-mode kernels (pb_render_d_vjp)."""
    global _pending
    if not isinstance(loss, Float32):
        raise TypeError("backward() expects a Float32")
    # collect the graph
    nodes, stack, seen = [], [loss], set()
    while stack:
        n = stack.pop()
        if id(n) in seen:
            continue
        seen.add(id(n))
        nodes.append(n)
        if n._bw is not None:
            for p in n._bw[0]:
                if p._tracked() and id(p) not in seen:
                    stack.append(p)
    for n in nodes:
        n._adj = None
    loss._adj = np.ones(loss.v.shape, np.float64)
    done_records = set()
    for n in sorted(nodes, key=lambda a: -a._id):
        if n._bw is None:
            if n.requires_grad and n._adj is not None:
                n._grad = n._adj.copy() if n._grad is None else n._grad + n._adj
            continue
        parents, fn = n._bw
        if isinstance(fn, _RenderRecord):   # a channel of a renderD image (the three have the largest ids of their record): run the VJP once
            if id(fn) not in done_records:
                done_records.add(id(fn))
                fn.backward()
            continue
        if n.requires_grad and n._adj is not None:   # a copy that was itself marked (adam.py: u = type(x)(u); set_requires_gradient(u))
            n._grad = n._adj.copy() if n._grad is None else n._grad + n._adj
        if n._adj is None:
            continue
        for p, g in zip(parents, fn(n._adj)):
            if p._tracked():
                g = _unbroadcast(g, p.v.shape)
                p._adj = g if p._adj is None else p._adj + g
    if free_graph:
        for n in nodes:
            n._adj = None
            if not n.requires_grad:
                n._bw = None
    _pending = []


def gradient(y):
    """reverse mode: dLoss/dy after `backward`; forward mode: the tangent of `y` after `forward`"""
    if isinstance(y, Vector3f):
        if any(c._grad is not None for c in y._comps()):
            n = len(y)
            return Vector3f(np.stack([np.broadcast_to(np.zeros(1) if c._grad is None else c._grad, (n,)) for c in y._comps()], axis=1).astype(np.float32))
        return Vector3f(y.tangent_numpy())
    if y._grad is not None:
        return Float32(y._grad.astype(np.float32))
    return Float32(_tan(y.d, y.v))


def detach(x):
    if isinstance(x, Vector3f):
        return Vector3f(x.numpy())
    if isinstance(x, Matrix4f):
        return Matrix4f(x.v.copy())
    return Float32(x.v.copy())


def slices(x):
    return len(x)


def sqr(x):
    return x * x


def sqrt(x):
    x = x if isinstance(x, (Float32, Vector3f)) else Float32(x)
    if isinstance(x, Vector3f):
        return Vector3f._from_comps(sqrt(x.x), sqrt(x.y), sqrt(x.z))
    r = np.sqrt(x.v)
    return Float32._make(r, None if x.d is None else x.d / (2 * r), (x,), lambda g: (g / (2 * r),))


def abs(x):   # noqa: A001  (enoki.abs)
    if isinstance(x, Vector3f):
        return Vector3f._from_comps(abs(x.x), abs(x.y), abs(x.z))
    x = x if isinstance(x, Float32) else Float32(x)
    s = np.sign(x.v)
    return Float32._make(np.abs(x.v), None if x.d is None else x.d * s, (x,), lambda g: (g * s,))


def hsum(x):
    if isinstance(x, Vector3f):
        return hsum(x.x) + hsum(x.y) + hsum(x.z)
    x = x if isinstance(x, Float32) else Float32(x)
    n = x.v.shape
    return Float32._make(np.atleast_1d(x.v.astype(np.float64).sum()).astype(np.float32), None if x.d is None else np.atleast_1d(x.d.sum()), (x,),
                         lambda g: (np.broadcast_to(g, n),))


def hmean(x):
    if isinstance(x, Vector3f):
        return (hmean(x.x) + hmean(x.y) + hmean(x.z)) / 3.0
    x = x if isinstance(x, Float32) else Float32(x)
    return hsum(x) / float(len(x.v))


def dot(a, b):
    return a.x * b.x + a.y * b.y + a.z * b.z


def squared_norm(v):
    return v.x * v.x + v.y * v.y + v.z * v.z


def norm(v):
    return sqrt(squared_norm(v))


class _RenderRecord:
    """what `Integrator.renderD` leaves on the tape: the three channel arrays of the image and a callback that turns dL/dI (numpy,
    (H*W, 3)) into [(array, adjoint)] for the arrays stored in the scene (compat/psdr_cuda)"""

    def __init__(self, channels, vjp):
        self.channels, self.vjp = channels, vjp

    def backward(self):
        n = len(self.channels[0].v)
        dLdI = np.stack([np.zeros(n) if c._adj is None else np.broadcast_to(c._adj, (n,)) for c in self.channels], axis=1).astype(np.float32)
        for arr, g in self.vjp(dLdI):
            comps = _leaves(arr)
            g = np.asarray(g, np.float64).reshape(-1, len(comps))
            for k, c in enumerate(comps):
                gk = _unbroadcast(g[:, k], c.v.shape)
                c._adj = gk if c._adj is None else c._adj + gk


def _attach_render(image, inputs, vjp):
    """called by compat/psdr_cuda: put a renderD image (Vector3f) on the tape, depending on the scene arrays `inputs`"""
    parents = []
    for arr in inputs:
        parents += [c for c in _leaves(arr) if c._tracked()]
    if not parents:
        return
    rec = _RenderRecord(image._comps(), vjp)
    for c in image._comps():
        c._bw = (tuple(parents), rec)
