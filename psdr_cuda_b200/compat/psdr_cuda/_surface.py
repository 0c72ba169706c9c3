"""The small value types and utilities of the reference's Python module (src/psdr.cpp:75-164, 242-265) that are not part of the rendering
path: rays, frames, sample records, the discrete / hyper-cube distributions, Bitmap.eval / load_openexr, and the Mesh helpers
(vertex_normals, edge_indices, sample_position). In the reference they are Enoki CUDA arrays; here they are host-side numpy mirrors with the
same names, argument meaning and results, so scripts that poke at a scene keep working. renderC / renderD never call any of this — the hot
path is the CUDA library behind the C ABI. Each is checked against the reference's own code compiled for the CPU
(tests/test_host_module.py)."""
import numpy as np

f32 = np.float32


def _v(a, n):
    a = np.asarray(a, f32)
    return a.reshape(-1, n) if a.ndim != 2 else a


class _Ray:
    """psdr::Ray (include/psdr/core/ray.h:9-30): o, d as (n, 3) arrays"""

    def __init__(self, o=None, d=None):
        self.o = None if o is None else _v(o, 3)
        self.d = None if d is None else _v(d, 3)

    def reversed(self):
        return type(self)(self.o, -self.d)


class RayC(_Ray):
    pass


class RayD(_Ray):
    pass


def coordinate_system(n):
    """frame.h:9-30 (Duff et al. 2017): s, t for unit normals n (m, 3)"""
    n = _v(n, 3)
    sign = np.copysign(f32(1), n[:, 2])
    a = -f32(1) / (sign + n[:, 2])
    b = n[:, 0] * n[:, 1] * a
    s = np.stack([np.copysign(n[:, 0] * n[:, 0] * a, n[:, 2]) + f32(1), np.copysign(b, n[:, 2]), np.where(np.signbit(n[:, 2]), n[:, 0], -n[:, 0])], axis=1)
    t = np.stack([b, sign + n[:, 1] * n[:, 1] * a, -n[:, 1]], axis=1)
    return s.astype(f32), t.astype(f32)


class _Frame:
    """psdr::Frame (frame.h:33-52)"""

    def __init__(self, n=None):
        self.s = self.t = self.n = None
        if n is not None:
            self.n = _v(n, 3)
            self.s, self.t = coordinate_system(self.n)

    def to_local(self, v):
        v = _v(v, 3)
        return np.stack([(v * self.s).sum(1), (v * self.t).sum(1), (v * self.n).sum(1)], axis=1)

    def to_world(self, v):
        v = _v(v, 3)
        return self.s * v[:, :1] + self.t * v[:, 1:2] + self.n * v[:, 2:3]


class FrameC(_Frame):
    pass


class FrameD(_Frame):
    pass


class _SampleRecord:
    """records.h:11-20"""

    def __init__(self, pdf=None, is_valid=None):
        self.pdf, self.is_valid = pdf, is_valid


class SampleRecordC(_SampleRecord):
    pass


class SampleRecordD(_SampleRecord):
    pass


class _PositionSample(_SampleRecord):
    """records.h:23-34"""

    def __init__(self, p=None, n=None, J=None, pdf=None, is_valid=None):
        super().__init__(pdf, is_valid)
        self.p, self.n, self.J = p, n, J


class PositionSampleC(_PositionSample, SampleRecordC):
    pass


class PositionSampleD(_PositionSample, SampleRecordD):
    pass


class BoundarySegSampleDirect(SampleRecordC):
    """records.h:35-45: what Scene.sample_boundary_segment_direct returns"""

    def __init__(self, p0=None, edge=None, edge2=None, p2=None, n=None, pdf=None, is_valid=None):
        super().__init__(pdf, is_valid)
        self.p0, self.edge, self.edge2, self.p2, self.n = p0, edge, edge2, p2, n


class DiscreteDistribution:
    """src/core/pmf.cpp:7-50: inclusive fp32 prefix sums; sample = first index whose cmf is not below u * sum"""

    def __init__(self):
        self.m_size = 0
        self.sum = f32(0)
        self._pmf = self._cmf = self._pmf_normalized = None

    def init(self, pmf):
        pmf = np.asarray(pmf, f32).reshape(-1)
        self.m_size = len(pmf)
        cmf = np.empty_like(pmf)
        acc = f32(0)
        for i, x in enumerate(pmf):          # sequential fp32, the order a prefix sum defines
            acc = f32(acc + x)
            cmf[i] = acc
        self._pmf, self._cmf, self.sum = pmf, cmf, acc
        self._pmf_normalized = pmf / acc

    def pmf(self):
        return self._pmf_normalized

    def _search(self, x):
        idx = np.searchsorted(self._cmf, x, side="left")      # first i with cmf[i] >= x
        return np.minimum(idx, self.m_size - 1).astype(np.int32)

    def sample(self, samples):
        samples = np.asarray(samples, f32).reshape(-1)
        if self.m_size == 1:
            return np.zeros(1, np.int32), np.ones(1, f32)
        idx = self._search(samples * self.sum)
        return idx, self._pmf[idx] / self.sum

    def sample_reuse(self, samples):
        """-> (idx, pdf); `samples` (float32 array) is rescaled in place for reuse, as the reference's reference argument is"""
        if self.m_size == 1:
            return np.zeros(1, np.int32), np.ones(1, f32)
        samples *= self.sum
        idx = self._search(samples)
        samples -= np.where(idx > 0, self._cmf[np.maximum(idx - 1, 0)], f32(0))
        pmf = self._pmf[idx]
        np.divide(samples, pmf, out=samples, where=pmf > 0)
        np.clip(samples, 0, 1, out=samples)
        return idx, pmf / self.sum


class _HyperCubeDistribution:
    """src/core/cube_distrb.cpp:8-62 (the last dimension runs fastest)"""
    ndim = 0

    def __init__(self):
        self.m_ready = False
        self.m_resolution = np.zeros(self.ndim, np.int32)
        self.m_distrb = DiscreteDistribution()
        self.m_num_cells = 0
        self.cells = None
        self.m_unit = None

    def set_resolution(self, reso):
        reso = np.asarray(reso, np.int32).reshape(-1)
        assert len(reso) == self.ndim
        if np.array_equal(reso, self.m_resolution):
            return
        self.m_num_cells = int(np.prod(reso.astype(np.int64)))
        self.m_resolution = reso.copy()
        self.m_unit = (f32(1) / reso.astype(f32)).astype(f32)
        cur = np.arange(self.m_num_cells, dtype=np.int32)
        cells = np.zeros((self.m_num_cells, self.ndim), np.int32)
        for i in range(self.ndim - 1):
            denom = int(np.prod(reso[i + 1:].astype(np.int64)))
            cells[:, i] = cur // denom
            cur = cur - cells[:, i] * denom
        cells[:, self.ndim - 1] = cur
        self.cells = cells
        self.m_ready = False

    def set_mass(self, pmf):
        pmf = np.asarray(pmf, f32).reshape(-1)
        if len(pmf) != self.m_num_cells:
            raise RuntimeError("static_cast<int>(slices(pmf)) == m_num_cells")
        self.m_distrb.init(pmf)
        self.m_ready = True

    def sample_reuse(self, samples):
        """samples: float32 (m, ndim), warped in place -> pdf (m,)"""
        if not self.m_ready:
            raise RuntimeError("m_ready")
        last = samples[:, self.ndim - 1].copy()
        idx, pdf = self.m_distrb.sample_reuse(last)
        samples[:, self.ndim - 1] = last
        samples += self.cells[idx].astype(f32)
        samples *= self.m_unit
        return (pdf * f32(self.m_num_cells)).astype(f32)

    def pdf(self, p):
        if not self.m_ready:
            raise RuntimeError("m_ready")
        p = _v(p, self.ndim)
        ip = np.floor(p * self.m_resolution.astype(f32)).astype(np.int64)
        valid = np.all((ip >= 0) & (ip < self.m_resolution), axis=1)
        idx = ip[:, 0]
        for i in range(1, self.ndim):
            idx = idx * int(self.m_resolution[i]) + ip[:, i]
        idx = np.where(valid, idx, 0)
        return np.where(valid, self.m_distrb.pmf()[idx] * f32(self.m_num_cells), f32(0)).astype(f32)


class HyperCubeDistribution2f(_HyperCubeDistribution):
    ndim = 2


class HyperCubeDistribution3f(_HyperCubeDistribution):
    ndim = 3


# ---- Bitmap (src/core/bitmap.cpp) -------------------------------------------------------------------------------------------------------------
def bitmap_eval(bitmap, uv, flip_v=True):
    """Bitmap::eval (bitmap.cpp:56-96): bilinear lookup, v flipped by default, wrap by uv - floor(uv), the last texel clamped"""
    width, height = bitmap.resolution
    ch = bitmap.channels
    data = np.asarray(bitmap.data, f32).reshape(height * width, ch)
    uv = _v(uv, 2).copy()
    if width == 1 and height == 1:
        out = np.broadcast_to(data[0], (len(uv), ch)).copy()
        return out[:, 0] if ch == 1 else out
    if width < 2 or height < 2:
        raise RuntimeError("Bitmap: invalid resolution!")
    if flip_v:
        uv[:, 1] = -uv[:, 1]
    uv -= np.floor(uv)
    uv *= np.array([width - 1, height - 1], f32)
    pos = np.floor(uv).astype(np.int32)
    w1 = uv - pos.astype(f32)
    w0 = f32(1) - w1
    pos = np.minimum(pos, np.array([width - 2, height - 2], np.int32))
    idx = pos[:, 1] * width + pos[:, 0]
    v00, v10, v01, v11 = data[idx], data[idx + 1], data[idx + width], data[idx + width + 1]
    v0 = w0[:, :1] * v00 + w1[:, :1] * v10
    v1 = w0[:, :1] * v01 + w1[:, :1] * v11
    out = (w0[:, 1:2] * v0 + w1[:, 1:2] * v1).astype(f32)
    return out[:, 0] if ch == 1 else out


# ---- Mesh helpers (src/shape/mesh.cpp) ----------------------------------------------------------------------------------------------------------
def _triangles(verts, faces):
    p0 = verts[faces[:, 0]]
    e1, e2 = verts[faces[:, 1]] - p0, verts[faces[:, 2]] - p0
    fn = np.cross(e1, e2).astype(f32)
    area = np.sqrt((fn * fn).sum(1)).astype(f32)
    return p0, e1, e2, fn, area


def mesh_vertex_normals(mesh):
    """Mesh.vertex_normals: object-space, area-weighted (mesh.cpp:19-51 applied to vertex_positions, mesh.cpp:219)"""
    verts, faces = np.asarray(mesh.vertex_positions, f32), np.asarray(mesh.face_indices, np.int64)
    _, _, _, fn, area = _triangles(verts, faces)
    vn = np.zeros_like(verts)
    w = np.zeros(len(verts), f32)
    for i in range(3):
        np.add.at(vn, faces[:, i], fn)
        np.add.at(w, faces[:, i], area)
    vn = vn / w[:, None]
    return (vn / np.sqrt((vn * vn).sum(1, keepdims=True))).astype(f32)


def mesh_edge_indices(mesh):
    """Mesh.edge_indices(): (4, ne) int32 = end points, then the one or two faces sharing the edge (-1 on a boundary), in the order of the
    reference's std::map keyed by (min vertex, max vertex) (mesh.cpp:143-203; psdr.cpp:263 returns head<4>)"""
    faces = np.asarray(mesh.face_indices, np.int64)
    if not mesh.enable_edges or len(faces) == 0:
        return np.zeros((4, 0), np.int32)
    nf = len(faces)
    a = faces.reshape(-1)                                  # corner k of face f at 3 f + k
    b = faces[:, [1, 2, 0]].reshape(-1)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    f = np.repeat(np.arange(nf, dtype=np.int64), 3)
    order = np.lexsort((np.arange(3 * nf), hi, lo))        # map order; ties keep the order of insertion (face, corner)
    lo, hi, f = lo[order], hi[order], f[order]
    first = np.ones(len(lo), bool)
    first[1:] = (lo[1:] != lo[:-1]) | (hi[1:] != hi[:-1])
    start = np.nonzero(first)[0]
    count = np.diff(np.append(start, len(lo)))
    if (count > 2).any():
        raise RuntimeError("Edge shared by more than 2 faces")
    f0 = f[start]
    f1 = np.where(count == 2, f[np.minimum(start + 1, len(f) - 1)], -1)
    if ((count == 2) & (f0 == f1)).any():
        raise RuntimeError("Duplicated faces")
    return np.stack([lo[start], hi[start], f0, f1]).astype(np.int32)


def mesh_sample_position(mesh, sample2, active=True):
    """Mesh.sample_position (mesh.cpp:277-303): a face by area (sample_reuse on x), then a uniform point on it; world space"""
    verts = np.asarray(mesh.vertex_positions, f32)
    M = np.asarray(mesh.to_world, f32)
    h = np.concatenate([verts, np.ones((len(verts), 1), f32)], axis=1) @ M.T
    world = (h[:, :3] / h[:, 3:4]).astype(f32)
    faces = np.asarray(mesh.face_indices, np.int64)
    p0, e1, e2, fn, area2 = _triangles(world, faces)
    distrb = DiscreteDistribution()
    distrb.init(area2 * f32(0.5))
    s = _v(sample2, 2).copy()
    x = s[:, 0].copy()
    idx, _ = distrb.sample_reuse(x)
    if distrb.m_size == 1:
        idx = np.zeros(len(s), np.int32)
    s[:, 0] = x
    t = np.sqrt(np.maximum(f32(1) - s[:, 0], 0))           # warp::square_to_uniform_triangle (warp.h:77-80)
    u, v = f32(1) - t, t * s[:, 1]
    p = (e1[idx] * u[:, None] + (e2[idx] * v[:, None] + p0[idx])).astype(f32)
    n = (fn[idx] / area2[idx][:, None]).astype(f32)
    total = f32((area2 * f32(0.5)).sum(dtype=f32))
    return PositionSampleC(p=p, n=n, J=np.ones(len(s), f32), pdf=np.full(len(s), f32(1) / total, f32), is_valid=np.broadcast_to(np.asarray(active, bool), (len(s),)).copy())
