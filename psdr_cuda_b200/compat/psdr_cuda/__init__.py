"""`import psdr_cuda` — the reference's module name (src/psdr.cpp:41) on top of the B200-native hot path.

The compiled host layer `_psdr_host` (pybind11, C++) owns scene ingest, the scene objects and the calls into the C ABI;
this file adds what is Python-side in this design: device results come back as torch CUDA tensors, and `renderD` is a
torch.autograd node whose backward is `pb_render_d_vjp` (the reference returns Enoki autodiff arrays and relies on
`ek.backward`; Enoki is not a dependency here).

    scene = psdr_cuda.Scene(); scene.load_file("scenes/cbox_bunny.xml")
    albedo = scene.parameter("BSDF[id=white]", "reflectance")        # torch leaf, requires_grad
    verts  = scene.parameter("Mesh[1]", "vertex_positions")
    scene.configure()
    img = psdr_cuda.PathIntegrator(5).renderD(scene, 0)                # (H*W, 3) CUDA tensor attached to the graph
    img.square().mean().backward(); albedo.grad, verts.grad

Put this directory's parent (psdr_cuda_b200/compat) on sys.path, or `import psdr_cuda_b200.compat` once.
"""
import os

import numpy as np

from . import _psdr_host as _h
from . import _surface
from ._psdr_host import (BSDF, AreaLight, BitmapD, DiffuseBSDF, Emitter, EnvironmentMap, Mesh, Object, PerspectiveCamera, RenderOption,  # noqa: F401
                         RoughConductorBSDF)
from ._surface import (BoundarySegSampleDirect, DiscreteDistribution, FrameC, FrameD, HyperCubeDistribution2f, HyperCubeDistribution3f, PositionSampleC,  # noqa: F401
                       PositionSampleD, RayC, RayD, SampleRecordC, SampleRecordD)



class _BitmapCtor:
    """constructors of src/psdr.cpp:102-106,112-116: (), (value), (width, height, data), (file name)"""
    _channels = 0

    def __init__(self, *args):
        BitmapD.__init__(self, self._channels)
        if len(args) == 1 and isinstance(args[0], str):
            self.load_openexr(args[0])
        elif len(args) == 1:
            self.data = np.broadcast_to(np.asarray(args[0], np.float32).reshape(-1), (self._channels,)).copy()
        elif len(args) == 3:
            self.resolution = (int(args[0]), int(args[1]))
            self.data = np.asarray(args[2], np.float32).reshape(int(args[0]) * int(args[1]), self._channels)
        elif args:
            raise TypeError("Bitmap%dfD(): expected (), (value), (width, height, data) or (file name)" % self._channels)


class Bitmap1fD(_BitmapCtor, BitmapD):
    _channels = 1


class Bitmap3fD(_BitmapCtor, BitmapD):
    _channels = 3


_envmap_init = EnvironmentMap.__init__


def _envmap_from_file(self, file_name):
    """EnvironmentMap(file name) (envmap.h:14-16): a latitude-longitude OpenEXR radiance map"""
    _envmap_init(self)
    img = _read_exr(file_name, 3)
    self.radiance.resolution = (img.shape[1], img.shape[0])
    self.radiance.data = img.reshape(-1, 3)


EnvironmentMap.__init__ = _envmap_from_file
Sensor = PerspectiveCamera   # src/psdr.cpp:216-222: the reference's only Sensor is the PerspectiveCamera; one class serves both names


def _read_exr(path, channels):
    """EXR decode for textures / environment maps (the reference uses tinyexr, src/core/bitmap_loader.cpp:13-53)"""
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise RuntimeError("Failed to load EXR: " + path)
    img = img.astype(np.float32)
    if img.ndim == 2:
        img = img[:, :, None]
    else:
        img = img[:, :, [2, 1, 0] + list(range(3, img.shape[2]))]
    return np.ascontiguousarray(img[:, :, :channels])


def _bitmap_load_openexr(self, file_name):
    """Bitmap::load_openexr (bitmap.cpp:34-43): channel 0 for Bitmap1fD, RGB for Bitmap3fD"""
    img = _read_exr(file_name, self.channels)
    self.resolution = (img.shape[1], img.shape[0])
    self.data = img.reshape(-1, self.channels) if self.channels == 3 else img.reshape(-1)


BitmapD.load_openexr = _bitmap_load_openexr
BitmapD.eval = _surface.bitmap_eval                                               # src/psdr.cpp:107,117
Mesh.vertex_normals = property(_surface.mesh_vertex_normals)                      # src/psdr.cpp:253
Mesh.edge_indices = _surface.mesh_edge_indices                                    # src/psdr.cpp:263
Mesh.sample_position = _surface.mesh_sample_position                              # src/psdr.cpp:248-249
Mesh.bsdf = property(lambda self: None)                                           # src/psdr.cpp:251; param_map entries resolve it (see _Proxy)


def _mesh_configure(self):
    """Mesh::configure (mesh.cpp:215-274) for a stand-alone mesh: validates it; the tables are built on the device by Scene.configure"""
    if self.num_vertices == 0 or self.num_faces == 0:
        raise RuntimeError("Mesh::configure: empty mesh")
    _surface.mesh_vertex_normals(self)


Mesh.configure = _mesh_configure                                                  # src/psdr.cpp:245


def _enoki():
    """the Enoki stand-in (psdr_cuda_b200/compat/enoki) if the caller imported it — examples/ do — else None"""
    import sys
    m = sys.modules.get("enoki")
    return m if m is not None and getattr(m, "__psdr_b200_shim__", False) else None


_TEXTURE_FIELDS = ("reflectance", "alpha_u", "alpha_v", "eta", "k", "specular_reflectance")


def _canonical_key(obj):
    """the param_map key the gradient layout uses for this object ("BSDF[i]" / "Mesh[i]"); other objects keep their type name"""
    t = obj.type_name()
    if t == "Mesh":
        return "Mesh[%d]" % obj.index
    if t in ("Diffuse", "RoughConductor"):
        return "BSDF[%d]" % obj.index
    return t


class _BitmapProxy:
    """a texture of a param_map entry (src/psdr.cpp:100-118): `.data` reads / accepts arrays of the Enoki stand-in, so that
    `ek.set_requires_gradient(bsdf.reflectance.data)` and `bsdf.reflectance.data = u` (examples/utils/adam.py:26-49) reach the scene"""

    def __init__(self, scene, key, field, bitmap):
        object.__setattr__(self, "_scene", scene)
        object.__setattr__(self, "_key", key)
        object.__setattr__(self, "_field", field)
        object.__setattr__(self, "_bm", bitmap)

    def __getattr__(self, name):
        bm = object.__getattribute__(self, "_bm")
        if name == "data" and _enoki() is not None:
            return self._scene._ek_array(self._key, self._field, lambda: self._to_ek(np.asarray(bm.data)))
        return getattr(bm, name)

    def _to_ek(self, a):
        ek = _enoki()
        return ek.Vector3f(a.reshape(-1, 3)) if self._bm.channels == 3 else ek.Float32(a.reshape(-1))

    def __setattr__(self, name, value):
        bm, scene = self._bm, self._scene
        if name == "data" and hasattr(value, "_tracked"):
            a = value.numpy()
            n = max(1, bm.resolution[0] * bm.resolution[1])
            a = np.broadcast_to(a.reshape(-1, bm.channels) if bm.channels == 3 else a.reshape(-1, 1), (n, bm.channels))
            bm.data = np.ascontiguousarray(a if bm.channels == 3 else a.reshape(-1), dtype=np.float32)
            scene._ek_inputs[(self._key, self._field)] = value
            t = value.tangent_numpy() if hasattr(value, "tangent_numpy") else (None if value.d is None else value.d)
            has_t = value.has_tangent() if hasattr(value, "has_tangent") else value.d is not None
            scene._fwd_texture[(self._key, self._field)] = np.broadcast_to(np.asarray(t, np.float32).reshape(-1, bm.channels), (n, bm.channels)).reshape(-1).copy() if has_t else None
            if has_t or value._tracked():
                bm.requires_grad = True
            return
        setattr(bm, name, value)

    def __repr__(self):
        return repr(self._bm)


class _Proxy:
    """param_map entry: forwards to the C++ object and accepts the Enoki stand-in's arrays (value + tangent) where the
    reference accepts Enoki autodiff arrays (src/psdr.cpp:242-265)"""

    def __init__(self, scene, key, obj):
        object.__setattr__(self, "_scene", scene)
        object.__setattr__(self, "_key", key)
        object.__setattr__(self, "_obj", obj)

    def __getattr__(self, name):
        obj = object.__getattribute__(self, "_obj")
        if name == "vertex_positions" and obj.type_name() == "Mesh":
            scene = object.__getattribute__(self, "_scene")
            if obj.index in scene._device_newer:               # last set from a CUDA leaf: bring the host mirror up to date
                scene._get_vertices_device(obj.index)
                scene._device_newer.discard(obj.index)
        if name == "vertex_positions" and _enoki() is not None:   # one array object per mesh, so that set_requires_gradient on it sticks
            scene = object.__getattribute__(self, "_scene")
            return scene._ek_array(_canonical_key(obj), "vertex_positions", lambda: _enoki().Vector3f(obj.vertex_positions))
        if name in _TEXTURE_FIELDS and obj.type_name() in ("Diffuse", "RoughConductor") and hasattr(obj, name):
            return _BitmapProxy(object.__getattribute__(self, "_scene"), _canonical_key(obj), name, getattr(obj, name))
        if name == "bsdf" and obj.type_name() == "Mesh":      # Mesh::m_bsdf (src/psdr.cpp:251)
            scene = object.__getattribute__(self, "_scene")
            return scene._raw_param_map().get("BSDF[%d]" % obj.bsdf_index) if obj.bsdf_index >= 0 else None
        return getattr(obj, name)

    def __setattr__(self, name, value):
        obj, scene = self._obj, self._scene
        if name == "vertex_positions" and hasattr(value, "tangent_numpy"):
            obj.vertex_positions = value.numpy()
            scene._fwd_vertex[obj_index(obj)] = value.tangent_numpy() if value.has_tangent() else None
            scene._ek_inputs[(_canonical_key(obj), "vertex_positions")] = value
            if value.has_tangent() or value._tracked():
                obj.requires_grad = True
            return
        setattr(obj, name, value)

    def set_transform(self, mat, set_left=True):
        obj, scene = self._obj, self._scene
        if hasattr(mat, "v") and hasattr(mat, "d"):          # Matrix4f of the Enoki stand-in
            if obj.type_name() == "Mesh":
                obj.set_transform(mat.v, set_left)
            else:
                obj.set_transform(mat.v)
            if obj.type_name() == "Mesh":
                scene._fwd_transform[(obj_index(obj), bool(set_left))] = mat.d
                if mat.d is not None:
                    obj.requires_grad = True
            elif isinstance(obj, EnvironmentMap):        # examples/utils/differential.py envmap_rotate
                scene._fwd_envmap = None if mat.d is None else np.asarray(mat.d, np.float32).reshape(16).copy()
                if mat.d is not None:
                    obj.transform_requires_grad = True
            elif mat.d is not None:
                raise RuntimeError("derivatives w.r.t. this object's transform are not implemented yet")
        else:
            obj.set_transform(np.asarray(mat, dtype=np.float32), set_left) if obj.type_name() == "Mesh" else obj.set_transform(np.asarray(mat, dtype=np.float32))

    def __repr__(self):
        obj = self._obj
        if obj.type_name() == "Mesh":   # Mesh::to_string (mesh.cpp:421-427)
            b = self.bsdf
            return "Mesh[nv=%d, nf=%d%s, bsdf=%s]" % (obj.num_vertices, obj.num_faces, ", id=" + obj.id if obj.id else "", repr(b) if b is not None else "None")
        return repr(obj)


def obj_index(obj):
    return obj.index


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("psdr_cuda needs a CUDA device: there is no CPU fallback")
    return torch


class Scene(_h.Scene):
    """Scene (src/psdr.cpp:269-278) + a registry of differentiable parameters held as torch leaves."""

    def __init__(self, device=0):
        super().__init__(device)
        self._device = device
        self._params = {}   # (key, field) -> torch leaf
        self._fwd_vertex = {}      # mesh index -> (nv, 3) object-space vertex tangent (Enoki stand-in, forward mode)
        self._fwd_transform = {}   # (mesh index, left?) -> 4x4 tangent of the transform
        self._fwd_texture = {}     # (BSDF key, field) -> flat texel tangent
        self._fwd_envmap = None    # 16 floats: tangent of the matrix EnvironmentMap.set_transform sets
        self._ek_inputs = {}       # (canonical key, field) -> array of the Enoki stand-in stored in / read from the scene
        self._device_newer = set() # meshes whose vertices were last set from a CUDA leaf (the host mirror is refreshed when it is read)

    @property
    def param_map(self):
        raw = _h.Scene.param_map.fget(self)
        return {k: _Proxy(self, k, v) for k, v in raw.items()}

    def _raw_param_map(self):
        return _h.Scene.param_map.fget(self)

    def _ek_array(self, key, field, make):
        """the one array object of the Enoki stand-in that stands for `param_map[key].<field>`"""
        arr = self._ek_inputs.get((key, field))
        if arr is None:
            arr = self._ek_inputs[(key, field)] = make()
        return arr

    def _push_ek_flags(self):
        """arrays marked by ek.set_requires_gradient after they were read from the scene (docs/inverse_diff_render.rst:66) -> requires_grad"""
        raw = self._raw_param_map()
        for (key, field), arr in self._ek_inputs.items():
            if not arr._tracked() or key not in raw:
                continue
            if field == "vertex_positions":
                raw[key].requires_grad = True
            else:
                getattr(raw[key], field).requires_grad = True

    def _forward_tangent(self):
        """flat tangent vector (layout of the gradient vector) from what the Enoki stand-in stored on the scene: explicit vertex
        tangents plus mesh-transform tangents mapped to object-space vertex tangents: u = M^-1 (dL raw R + L raw dR) (x, 1)"""
        torch = _torch()
        flat = np.zeros(max(1, self.grad_size()), np.float32)
        raw = self._raw_param_map()
        for key, field, off, cnt in self.grad_layout():
            if field != "vertex_positions":
                continue
            m = raw[key]
            i = m.index
            u = np.zeros((m.num_vertices, 3), np.float64)
            if self._fwd_vertex.get(i) is not None:
                u += self._fwd_vertex[i]
            dL, dR = self._fwd_transform.get((i, True)), self._fwd_transform.get((i, False))
            if dL is not None or dR is not None:
                L, R, W = m.to_world_left.astype(np.float64), m.to_world_right.astype(np.float64), m.to_world_raw.astype(np.float64)
                M = L @ W @ R
                dM = (0 if dL is None else dL.astype(np.float64) @ W @ R) + (0 if dR is None else L @ W @ dR.astype(np.float64))
                x = np.concatenate([m.vertex_positions.astype(np.float64), np.ones((m.num_vertices, 1))], axis=1)
                dworld = (x @ dM.T)[:, :3]
                u += np.linalg.solve(M[:3, :3], dworld.T).T
            flat[off:off + cnt] = u.reshape(-1).astype(np.float32)
        for key, field, off, cnt in self.grad_layout():      # texture tangents (examples/utils/differential.py material_roughness)
            t = self._fwd_texture.get((key, field))
            if t is not None and t.size == cnt:
                flat[off:off + cnt] = t
            if field == "to_world_left" and key.startswith("Emitter") and self._fwd_envmap is not None and cnt == 16:
                flat[off:off + cnt] = self._fwd_envmap
        return torch.from_numpy(flat).to("cuda:%d" % self._device)

    def parameter(self, key, field, requires_grad=True):
        """torch leaf mirroring `param_map[key].<field>` (e.g. ("BSDF[id=white]", "reflectance"), ("Mesh[1]", "vertex_positions")).
        Edit it in place / through an optimiser; `configure()` pushes the current value into the scene."""
        torch = _torch()
        obj = self._raw_param_map()[key]
        if field in ("to_world_left", "to_world_right"):   # Mesh.set_transform / append_transform leaves (src/psdr.cpp:246-247)
            value = getattr(obj, field)
        elif field == "scale":                              # EnvironmentMap.scale (src/psdr.cpp:237)
            value = np.float32(obj.scale)
        elif field == "to_world":                           # Sensor.to_world (src/psdr.cpp:220-224)
            value = obj.to_world
        elif field == "vertex_uv":                          # Mesh.vertex_uv (src/psdr.cpp:254)
            value = obj.vertex_uv
        else:
            value = obj.vertex_positions if field == "vertex_positions" else getattr(obj, field).data
        t = torch.tensor(np.asarray(value), dtype=torch.float32, device="cuda:%d" % self._device, requires_grad=requires_grad)
        self._params[(key, field)] = t
        return t

    _TEX_SLOT = {"reflectance": 0, "alpha_u": 1, "alpha_v": 2, "eta": 3, "k": 4, "specular_reflectance": 5}

    def configure(self):
        self._bind_stream()
        on_device = self.uploaded      # after the first configure the leaves update the scene device-to-device: an optimisation loop
        for (key, field), t in self._params.items():   # (torch.optim on CUDA leaves -> configure -> renderD -> backward) never touches the host
            obj = self._raw_param_map()[key]
            if on_device and field == "vertex_positions":
                self._set_vertices_device(obj.index, t.detach().contiguous().data_ptr())
                self._device_newer.add(obj.index)
                obj.requires_grad = bool(t.requires_grad) or obj.requires_grad
                continue
            if on_device and field in self._TEX_SLOT and obj.type_name() in ("Diffuse", "RoughConductor"):
                bm = getattr(obj, field)
                if t.numel() == max(1, bm.resolution[0] * bm.resolution[1]) * bm.channels:
                    self._set_texture_device(obj.index, self._TEX_SLOT[field], t.detach().contiguous().data_ptr())
                    bm.requires_grad = bool(t.requires_grad)
                    continue
            val = t.detach().cpu().numpy()
            if field == "vertex_positions":
                obj.vertex_positions = val
                obj.requires_grad = bool(t.requires_grad) or obj.requires_grad
            elif field == "scale":
                obj.scale = float(val)
                obj.scale_requires_grad = bool(t.requires_grad)
            elif field == "to_world":
                obj.to_world = val.astype(np.float32)
                obj.requires_grad = bool(t.requires_grad)
            elif field == "vertex_uv":
                obj.vertex_uv = val.astype(np.float32)
                obj.uv_requires_grad = bool(t.requires_grad)
            elif field == "to_world_left" and isinstance(obj, EnvironmentMap):   # EnvironmentMap.set_transform (src/psdr.cpp:238)
                obj.set_transform(val.astype(np.float32))
                obj.transform_requires_grad = bool(t.requires_grad)
            elif field in ("to_world_left", "to_world_right"):
                obj.set_transform(val.astype(np.float32), field == "to_world_left")
                obj.requires_grad = bool(t.requires_grad) or obj.requires_grad   # the transform gradient is a contraction of the vertex gradient
            else:
                bm = getattr(obj, field)
                bm.data = val
                bm.requires_grad = bool(t.requires_grad)
        if _enoki() is not None:
            self._push_ek_flags()
        self._bind_stream()
        super().configure()

    def _bind_stream(self):
        """run on torch's current stream: images, dL/dI and gradient vectors are torch tensors produced / consumed on it"""
        try:
            import torch
            if torch.cuda.is_available() and self._device >= 0:
                cur = int(torch.cuda.current_stream(self._device).cuda_stream)
                if getattr(self, "_stream", None) != cur:
                    self.set_stream(cur)
                    self._stream = cur
        except ImportError:
            pass

    def init_distributed(self, group=None, mode="pixels", tile_rows=4):
        """one process per GPU under torch.distributed: create the library's NCCL communicator for the group's ranks (the unique id
        travels through the group), shard the scene over them. From then on renderC / renderD return the complete film and
        backward() returns the complete gradient: one collective each, enqueued by the library on the scene's stream."""
        torch = _torch()
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = "cuda:%d" % self._device if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t = torch.frombuffer(bytearray(self.dist_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.dist_init(bytes(t.cpu().numpy().tobytes()), rank, world)
        self.set_shard_mode(mode, tile_rows)

    def sample_boundary_segment_direct(self, sample3, active=True):
        """Scene::sample_boundary_segment_direct (src/scene/scene.cpp:456-492, src/psdr.cpp:274): a point on a face edge and a point on
        an emitter per 3-D sample. `sample3`: (n, 3) array (numpy / torch / Vector3f of the Enoki stand-in); fields come back in kind."""
        torch = _torch()
        ek = _enoki()
        s3 = sample3.numpy() if hasattr(sample3, "tangent_numpy") else sample3
        s3 = torch.as_tensor(np.asarray(s3.detach().cpu()) if hasattr(s3, "detach") else np.asarray(s3), dtype=torch.float32).reshape(-1, 3).to("cuda:%d" % self._device)
        self._bind_stream()
        out = torch.empty((s3.shape[0], 17), dtype=torch.float32, device=s3.device)
        self._sample_boundary_segment_direct(int(s3.shape[0]), s3.data_ptr(), out.data_ptr())
        valid = out[:, 16] > 0
        if active is not True:
            act = torch.as_tensor(np.asarray(active), dtype=torch.bool, device=out.device).reshape(-1).expand(out.shape[0])
            valid = valid & act
        pdf = torch.where(valid, out[:, 15], torch.zeros_like(out[:, 15]))
        f = [out[:, 0:3], out[:, 3:6], out[:, 6:9], out[:, 9:12], out[:, 12:15]]
        if ek is not None or isinstance(sample3, np.ndarray):
            f = [x.cpu().numpy() for x in f]
            pdf, valid = pdf.cpu().numpy(), valid.cpu().numpy()
            if ek is not None:
                f = [ek.Vector3f(x) for x in f]
                pdf = ek.Float32(pdf)
        return BoundarySegSampleDirect(*f, pdf=pdf, is_valid=valid)

    def _leaves_in_layout_order(self):
        """registered leaves matched to the segments of the flat gradient vector"""
        canon = {}
        pm = self._raw_param_map()
        for (key, field), t in self._params.items():
            obj = pm[key]
            for k2, o2 in pm.items():
                if o2 is obj or (o2.type_name() == obj.type_name() and o2.id == obj.id and o2.id != "" and k2.split("[")[0] == key.split("[")[0]):
                    canon[(k2, field)] = t
            canon[(key, field)] = t
        out = []
        for key, field, off, cnt in self.grad_layout():
            out.append((canon.get((key, field)), off, cnt))
        return out

    def _transform_leaves(self):
        """registered mesh-transform leaves -> (tensor, mesh object, left?, offset, count of the mesh's vertex segment)"""
        pm = self._raw_param_map()
        seg = {}
        for key, field, off, cnt in self.grad_layout():
            if field == "vertex_positions":
                seg[pm[key].index] = (off, cnt)
        out = []
        for (key, field), t in self._params.items():
            if field in ("to_world_left", "to_world_right") and t.requires_grad and not isinstance(pm[key], EnvironmentMap) and pm[key].index in seg:
                out.append((t, pm[key], field == "to_world_left") + seg[pm[key].index])
        return out

    @staticmethod
    def _transform_gradient(torch, mesh, left, g_obj):
        """dL/d(to_world_left|right) from the object-space vertex gradient g_obj (nv, 3): world = M x with M = L W R (mesh.cpp:223),
        g_obj = M3^T g_world, so g_M[:3, :] = sum_v g_world_v (x_v, 1)^T, then g_L = g_M (W R)^T, g_R = (L W)^T g_M. Affine
        transforms only (the projective row gets no gradient), which is what examples/utils/differential.py builds."""
        dev, f64 = g_obj.device, torch.float64
        L = torch.tensor(mesh.to_world_left, dtype=f64, device=dev); W = torch.tensor(mesh.to_world_raw, dtype=f64, device=dev)
        R = torch.tensor(mesh.to_world_right, dtype=f64, device=dev)
        M = L @ W @ R
        x = torch.tensor(mesh.vertex_positions, dtype=f64, device=dev)
        x1 = torch.cat([x, torch.ones((x.shape[0], 1), dtype=f64, device=dev)], dim=1)
        g_world = g_obj.to(f64) @ torch.linalg.inv(M[:3, :3])
        g_M = torch.zeros((4, 4), dtype=f64, device=dev)
        g_M[:3, :] = g_world.T @ x1
        g = g_M @ (W @ R).T if left else (L @ W).T @ g_M
        return g.to(torch.float32)


def _image_tensor(scene):
    torch = _torch()
    return torch.empty((scene.opts.height * scene.opts.width, 3), dtype=torch.float32, device="cuda:%d" % scene._device)


class _IntegratorMixin:
    def renderC(self, scene, sensor_id=0):
        """Integrator.renderC (src/integrator/integrator.cpp:13-29) -> (H*W, 3) CUDA tensor, pixel = y*W + x"""
        scene._bind_stream()
        img = _image_tensor(scene)
        self._render_c(scene, sensor_id, img.data_ptr())
        if scene.shard_world > 1:
            scene._allreduce_image(img.data_ptr())     # the film of a sharded scene is complete when the call returns
        ek = _enoki()
        return ek.Vector3f(img.cpu().numpy()) if ek is not None else img

    def renderD(self, scene, sensor_id=0):
        """Integrator.renderD (src/integrator/integrator.cpp:32-60) -> image attached to torch.autograd; backward() runs the
        reverse-mode kernels (interior + boundary terms) and fills .grad of the registered parameters. With the Enoki stand-in
        imported the image is one of its arrays instead: ek.forward / ek.backward differentiate it (examples/run_test.py:126-129,
        docs/inverse_diff_render.rst:63-79)."""
        torch = _torch()
        ek = _enoki()
        scene._bind_stream()
        integ = self
        if ek is not None:
            img = _image_tensor(scene)
            self._render_d(scene, sensor_id, img.data_ptr())
            if scene.shard_world > 1:
                scene._allreduce_image(img.data_ptr())
            state = scene._render_d_state()
            out = ek.Vector3f(img.cpu().numpy())

            def run_forward():   # ek.forward(P): the derivative image is produced now
                dimg = torch.empty_like(img)
                scene._render_d_set_state(list(state))
                integ._render_d_jvp(scene, sensor_id, scene._forward_tangent().data_ptr(), dimg.data_ptr())
                if scene.shard_world > 1:
                    scene._allreduce_image(dimg.data_ptr())
                d = dimg.cpu().numpy()
                out.x.d, out.y.d, out.z.d = d[:, 0].copy(), d[:, 1].copy(), d[:, 2].copy()
            out._run_forward = run_forward
            ek._pending.append(out)

            inputs = [(key, field, arr) for (key, field), arr in scene._ek_inputs.items() if arr._tracked()]

            def vjp(dLdI):       # ek.backward(loss) reached this image
                g = torch.from_numpy(np.ascontiguousarray(dLdI, dtype=np.float32)).to(img.device)
                grad = torch.zeros(max(1, scene.grad_size()), dtype=torch.float32, device=img.device)
                scene._render_d_set_state(list(state))
                integ._render_d_vjp(scene, sensor_id, g.data_ptr(), grad.data_ptr())
                if scene.shard_world > 1:
                    scene._allreduce_grads(grad.data_ptr(), grad.numel())
                flat = grad.cpu().numpy()
                seg = {(key, field): (off, cnt) for key, field, off, cnt in scene.grad_layout()}
                return [(arr, flat[seg[(key, field)][0]:seg[(key, field)][0] + seg[(key, field)][1]]) for key, field, arr in inputs if (key, field) in seg]
            if inputs:
                ek._attach_render(out, [arr for _, _, arr in inputs], vjp)
            return out
        segs = scene._leaves_in_layout_order()
        xforms = scene._transform_leaves()
        leaves = [t for t, _, _ in segs if t is not None and t.requires_grad] + [x[0] for x in xforms]

        class _RenderD(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *inputs):
                img = _image_tensor(scene)
                integ._render_d(scene, sensor_id, img.data_ptr())
                if scene.shard_world > 1:
                    scene._allreduce_image(img.data_ptr())
                # the reference's tape differentiates each renderD image with its own samples, however many renders precede backward()
                # (one image per sensor in a multi-view loss): remember this render's sampler positions for the VJP
                ctx.render_state = scene._render_d_state()
                return img

            @staticmethod
            def backward(ctx, g):
                scene._bind_stream()
                grad = torch.zeros(max(1, scene.grad_size()), dtype=torch.float32, device=g.device)
                g = g.contiguous().float()
                scene._render_d_set_state(list(ctx.render_state))
                integ._render_d_vjp(scene, sensor_id, g.data_ptr(), grad.data_ptr())
                if scene.shard_world > 1:   # only a scene that was sharded explicitly exchanges anything: one sum of the gradient vector
                    scene._allreduce_grads(grad.data_ptr(), grad.numel())
                outs = []
                for t, off, cnt in segs:
                    if t is not None and t.requires_grad:
                        outs.append(grad[off:off + cnt].view_as(t))
                for t, mesh, left, off, cnt in xforms:
                    outs.append(Scene._transform_gradient(torch, mesh, left, grad[off:off + cnt].view(-1, 3)))
                return tuple(outs)

        return _RenderD.apply(*leaves) if leaves else _RenderD.apply()


    def forward(self, scene, tangents, sensor_id=0):
        """Forward mode (ek.forward(P) + ek.gradient(image) in examples/run_test.py:126-129): `tangents` maps registered
        parameters (the tensors returned by scene.parameter) to tangent tensors of the same shape; returns (image, d image).
        Unlisted parameters get a zero tangent."""
        torch = _torch()
        scene._bind_stream()
        img = _image_tensor(scene)
        self._render_d(scene, sensor_id, img.data_ptr())
        if scene.shard_world > 1:
            scene._allreduce_image(img.data_ptr())
        flat = torch.zeros(max(1, scene.grad_size()), dtype=torch.float32, device=img.device)
        for t, off, cnt in scene._leaves_in_layout_order():
            if t is None:
                continue
            for p, tan in tangents.items():
                if p is t:
                    flat[off:off + cnt] = tan.to(flat.device, torch.float32).reshape(-1)
        dimg = torch.empty_like(img)
        self._render_d_jvp(scene, sensor_id, flat.data_ptr(), dimg.data_ptr())
        if scene.shard_world > 1:
            scene._allreduce_image(dimg.data_ptr())
        return img, dimg


class DirectIntegrator(_IntegratorMixin, _h.DirectIntegrator):
    pass


class PathIntegrator(_IntegratorMixin, _h.PathIntegrator):
    pass


class FieldExtractionIntegrator(_IntegratorMixin, _h.FieldExtractionIntegrator):
    pass


Integrator = _h.Integrator
# src/psdr.cpp:283-286: renderC / renderD live on the Integrator base, so every integrator object answers them
Integrator.renderC, Integrator.renderD, Integrator.forward = _IntegratorMixin.renderC, _IntegratorMixin.renderD, _IntegratorMixin.forward
