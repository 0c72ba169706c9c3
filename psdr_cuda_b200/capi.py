"""ctypes binding of include/psdr_b200.h (lib/libpsdr_b200.so) — the lowest Python layer of the product.

torch is used only as the owner of device buffers (images, rays, gradients); every computation happens inside the
shared library. There is no CPU fallback: creating a context without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libpsdr_b200.so")
_lib = None

BSDF_DIFFUSE, BSDF_ROUGHCONDUCTOR = 0, 1
TEX = {"reflectance": 0, "alpha_u": 1, "alpha_v": 2, "eta": 3, "k": 4, "specular_reflectance": 5}
INTEG_DIRECT, INTEG_FIELD, INTEG_PATH = 0, 1, 2
FIELDS = {"silhouette": 0, "position": 1, "depth": 2, "geoNormal": 3, "shNormal": 4, "uv": 5}
MESH_FACE_NORMALS, MESH_ENABLE_EDGES = 1, 2
PARAM_BSDF_TEXTURE, PARAM_MESH_VERTICES, PARAM_ENVMAP_RADIANCE, PARAM_ENVMAP_SCALE, PARAM_SENSOR_TRANSFORM, PARAM_ENVMAP_TRANSFORM, PARAM_MESH_UV = 0, 1, 2, 3, 4, 5, 6

SYMBOLS = [
    "pb_ctx_create", "pb_ctx_destroy", "pb_last_error", "pb_version", "pb_ctx_set_batch", "pb_ctx_set_shard", "pb_ctx_set_stream", "pb_ctx_set_retain_limit",
    "pb_scene_set_options", "pb_scene_add_sensor", "pb_scene_set_sensor_transform", "pb_scene_add_bsdf", "pb_scene_set_bsdf_texture",
    "pb_scene_add_mesh", "pb_scene_set_mesh_vertices", "pb_scene_set_mesh_uvs", "pb_scene_set_mesh_transform", "pb_scene_add_area_emitter", "pb_scene_add_envmap", "pb_scene_set_envmap_radiance", "pb_scene_set_envmap_transform", "pb_scene_num_meshes", "pb_scene_configure",
    "pb_scene_reseed", "pb_scene_num_triangles", "pb_scene_get_triangle_info", "pb_scene_mesh_num_edges", "pb_scene_mesh_get_edges",
    "pb_trace", "pb_preprocess_secondary_edges", "pb_render_c", "pb_render_c_host", "pb_render_d", "pb_grad_require", "pb_grad_num_segments", "pb_grad_segment",
    "pb_grad_size", "pb_render_d_vjp", "pb_render_d_jvp", "pb_stats_launches", "pb_stats_last_trace_ms", "pb_stats_last_rays", "pb_stats_last_active_rays", "pb_ctx_set_bvh_refit", "pb_ctx_set_bvh_builder", "pb_stats_bvh", "pb_stats_last_trace_launches", "pb_stats_last_primary_ms",
    "pb_debug_set", "pb_debug_ray_buffer", "pb_debug_retained_rad", "pb_render_d_get_state", "pb_render_d_set_state",
    "pb_scene_set_mesh_vertices_device", "pb_scene_set_bsdf_texture_device", "pb_scene_get_mesh_vertices", "pb_scene_set_edge_importance",
    "pb_sample_boundary_segment_direct", "pb_scene_num_primary_edges", "pb_scene_get_primary_edges", "pb_scene_num_secondary_edges", "pb_scene_get_secondary_edges", "pb_ctx_set_shard_mode", "pb_dist_available", "pb_dist_unique_id", "pb_dist_init", "pb_dist_adopt_comm", "pb_dist_finalize", "pb_allreduce_grads", "pb_allreduce_image", "pb_stats_collectives",
]


class Integrator(C.Structure):
    _fields_ = [("kind", C.c_int), ("bsdf_samples", C.c_int), ("light_samples", C.c_int), ("hide_emitters", C.c_int),
                ("field", C.c_int), ("max_depth", C.c_int), ("use_guiding", C.c_int)]


def lib():
    """Load the shared library; fail loudly if it has not been built (python -m psdr_cuda_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError("psdr_cuda_b200: %s is missing — run `python psdr_cuda_b200/build.py` (no CPU fallback exists)" % _LIB_PATH)
        L = C.CDLL(_LIB_PATH)
        L.pb_last_error.restype = C.c_char_p
        L.pb_last_error.argtypes = [C.c_void_p]
        L.pb_grad_size.restype = C.c_int64
        L.pb_stats_launches.restype = C.c_int64
        L.pb_stats_collectives.restype = C.c_int64
        L.pb_stats_last_rays.restype = C.c_int64
        L.pb_stats_last_active_rays.restype = C.c_int64
        L.pb_stats_last_trace_ms.restype = C.c_float
        L.pb_stats_last_primary_ms.restype = C.c_float
        _lib = L
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _dp(t):
    """device pointer of a contiguous torch CUDA tensor"""
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def make_integrator(kind="direct", bsdf_samples=1, light_samples=1, hide_emitters=False, field="silhouette", max_depth=1, use_guiding=False):
    k = {"direct": INTEG_DIRECT, "field": INTEG_FIELD, "path": INTEG_PATH}[kind]
    return Integrator(k, bsdf_samples, light_samples, int(hide_emitters), FIELDS[field], max_depth, int(use_guiding))


class Context:
    def __init__(self, device=0):
        L = lib()
        h = C.c_void_p()
        if L.pb_ctx_create(int(device), C.byref(h)) != 0:
            raise RuntimeError(L.pb_last_error(None).decode())
        self.h = h
        self.device = int(device)
        self.width = self.height = 0
        self._stream = None     # None: the context's own stream (never ordered against torch's: see _bind_stream)
        self._stream_pinned = False

    def close(self):
        if self.h is not None:
            lib().pb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(lib().pb_last_error(self.h).decode())

    def _id(self, rc):
        if rc < 0:
            raise RuntimeError(lib().pb_last_error(self.h).decode())
        return rc

    # --- scene description -----------------------------------------------------------------------------------
    def set_options(self, width, height, spp, sppe=0, sppse=0):
        self._chk(lib().pb_scene_set_options(self.h, width, height, spp, sppe, sppse))
        self.width, self.height = width, height

    def set_batch(self, lanes):
        self._chk(lib().pb_ctx_set_batch(self.h, C.c_int64(lanes)))

    def set_retain_limit(self, nbytes):
        self._chk(lib().pb_ctx_set_retain_limit(self.h, C.c_int64(nbytes)))

    def set_stream(self, cuda_stream):
        """run on the caller's stream (int handle, e.g. torch.cuda.current_stream().cuda_stream; 0 = legacy default) from now on"""
        self._chk(lib().pb_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))
        self._stream, self._stream_pinned = int(cuda_stream), True

    def _bind_stream(self):
        """Every call that touches caller buffers (images, dL/dI, gradient vectors are torch tensors produced on torch's current
        stream) runs on that stream, so the library's kernels are ordered against their producers and consumers. Without this the
        context's private non-blocking stream would race with e.g. the zero-fill of a fresh gradient tensor."""
        if self._stream_pinned:
            return
        import torch
        cur = int(torch.cuda.current_stream(self.device).cuda_stream)
        if cur != self._stream:
            self._chk(lib().pb_ctx_set_stream(self.h, C.c_void_p(cur)))
            self._stream = cur

    def set_shard(self, rank, world):
        self._chk(lib().pb_ctx_set_shard(self.h, rank, world))

    def set_shard_mode(self, mode, tile_rows=0):
        """"samples" (every rank renders spp / world samples of every pixel) or "pixels" (image-row tiles of tile_rows rows, 0 = one block per rank)"""
        self._chk(lib().pb_ctx_set_shard_mode(self.h, {"samples": 0, "pixels": 1}.get(mode, mode), int(tile_rows)))

    # --- several GPUs: NCCL communicator owned by the library, collectives enqueued on the context's stream ----------------
    def dist_unique_id(self):
        buf = C.create_string_buffer(128)
        self._chk(lib().pb_dist_unique_id(self.h, buf))
        return buf.raw

    def dist_init(self, unique_id, rank, world):
        self._chk(lib().pb_dist_init(self.h, C.c_char_p(bytes(unique_id)), int(rank), int(world)))

    def dist_init_from_torch(self, group=None):
        """create the library's communicator for the ranks of an initialised torch.distributed group (the id travels through the group)"""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = "cuda:%d" % self.device if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t = torch.frombuffer(bytearray(self.dist_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.dist_init(bytes(t.cpu().numpy().tobytes()), rank, world)

    def dist_finalize(self):
        self._chk(lib().pb_dist_finalize(self.h))

    def allreduce_grads(self, grad):
        self._bind_stream()
        self._chk(lib().pb_allreduce_grads(self.h, _dp(grad), C.c_int64(grad.numel())))
        return grad

    def allreduce_image(self, img):
        self._bind_stream()
        self._chk(lib().pb_allreduce_image(self.h, _dp(img)))
        return img

    def add_sensor(self, fov, near, far, to_world):
        return self._id(lib().pb_scene_add_sensor(self.h, C.c_float(fov), C.c_float(near), C.c_float(far), _p(_f(to_world))))

    def add_bsdf(self, type_):
        return self._id(lib().pb_scene_add_bsdf(self.h, type_))

    def set_bsdf_texture(self, bsdf, slot, data):
        t = _f(data)
        assert t.ndim == 3
        self._chk(lib().pb_scene_set_bsdf_texture(self.h, bsdf, TEX[slot] if isinstance(slot, str) else slot, _p(t), t.shape[1], t.shape[0]))

    def add_mesh(self, verts, faces, uvs=None, uv_faces=None, face_normals=False, enable_edges=True, bsdf=-1, to_world=None):
        v, f = _f(verts), _i(faces)
        has_uv = uvs is not None
        u, uf = (_f(uvs), _i(uv_faces)) if has_uv else (None, None)
        flags = (MESH_FACE_NORMALS if face_normals else 0) | (MESH_ENABLE_EDGES if enable_edges else 0)
        tw = _f(to_world if to_world is not None else np.eye(4))
        return self._id(lib().pb_scene_add_mesh(self.h, len(v), len(f), _p(v), _p(f), len(u) if has_uv else 0, _p(u), _p(uf), flags, bsdf, _p(tw)))

    def set_mesh_vertices(self, mesh, verts):
        self._chk(lib().pb_scene_set_mesh_vertices(self.h, mesh, _p(_f(verts))))

    def set_mesh_vertices_device(self, mesh, verts):
        """vertices from a CUDA tensor (nv, 3): no host round trip (an optimiser with its parameters on the GPU)"""
        self._bind_stream()
        self._chk(lib().pb_scene_set_mesh_vertices_device(self.h, mesh, _dp(verts.contiguous().float())))

    def set_bsdf_texture_device(self, bsdf, slot, data):
        self._bind_stream()
        self._chk(lib().pb_scene_set_bsdf_texture_device(self.h, bsdf, TEX[slot] if isinstance(slot, str) else slot, _dp(data.contiguous().float())))

    def get_mesh_vertices(self, mesh, nv):
        out = np.zeros((nv, 3), np.float32)
        self._chk(lib().pb_scene_get_mesh_vertices(self.h, mesh, _p(out)))
        return out

    def set_edge_importance(self, mode):
        """"length" (the reference) or "dihedral" (length x exterior dihedral angle: scene.cpp:230-233, disabled there)"""
        self._chk(lib().pb_scene_set_edge_importance(self.h, {"length": 0, "dihedral": 1}.get(mode, mode)))

    def set_mesh_uvs(self, mesh, uvs):
        self._chk(lib().pb_scene_set_mesh_uvs(self.h, mesh, _p(_f(uvs))))

    def set_mesh_transform(self, mesh, mat, left=True):
        self._chk(lib().pb_scene_set_mesh_transform(self.h, mesh, _p(_f(mat)), int(left)))

    def add_area_emitter(self, mesh, radiance):
        return self._id(lib().pb_scene_add_area_emitter(self.h, mesh, _p(_f(radiance))))

    def add_envmap(self, radiance, scale=1.0, to_world=None):
        r = _f(radiance)
        assert r.ndim == 3 and r.shape[2] == 3
        tw = _f(to_world if to_world is not None else np.eye(4))
        return self._id(lib().pb_scene_add_envmap(self.h, r.shape[1], r.shape[0], _p(r), C.c_float(scale), _p(tw)))

    def set_envmap_radiance(self, radiance, scale=1.0):
        r = _f(radiance)
        self._chk(lib().pb_scene_set_envmap_radiance(self.h, _p(r), C.c_float(scale)))

    def set_envmap_transform(self, left):
        self._chk(lib().pb_scene_set_envmap_transform(self.h, _p(_f(left))))

    def configure(self, reseed=False):
        self._bind_stream()
        if reseed:
            lib().pb_scene_reseed(self.h)
        self._chk(lib().pb_scene_configure(self.h))

    def load_description(self, desc, opts=None):
        """Feed a scene-description dict (sensors / bsdfs / meshes / emitters / opts) through the C ABI."""
        o = dict(desc["opts"])
        if opts:
            o.update(opts)
        self.set_options(o["width"], o["height"], o["spp"], o.get("sppe", 0), o.get("sppse", 0))
        for s in desc["sensors"]:
            self.add_sensor(s["fov"], s["near"], s["far"], s["to_world"])
        for b in desc["bsdfs"]:
            bi = self.add_bsdf(b["type"])
            for k in TEX:
                if k in b:
                    self.set_bsdf_texture(bi, k, b[k])
        if desc.get("envmap") is not None:      # the reference loads the env emitter before the shapes (scene_loader.cpp:222-230)
            e = desc["envmap"]
            self.add_envmap(e["radiance"], e["scale"], e["to_world"])
        emitter_of = {e["mesh"]: e for e in desc["emitters"]}
        for mi, m in enumerate(desc["meshes"]):
            self.add_mesh(m["verts"], m["faces"], m.get("uvs"), m.get("uv_faces"), m["face_normals"], m["enable_edges"], m["bsdf"], m["to_world"])
            if mi in emitter_of:
                self.add_area_emitter(mi, emitter_of[mi]["radiance"])

    # --- inspection --------------------------------------------------------------------------------------------
    def triangle_info(self):
        n = lib().pb_scene_num_triangles(self.h)
        out = np.empty((n, 22), dtype=np.float32)
        self._chk(lib().pb_scene_get_triangle_info(self.h, _p(out)))
        return out

    def mesh_edges(self, mesh):
        n = self._id(lib().pb_scene_mesh_num_edges(self.h, mesh))
        out = np.empty((n, 5), dtype=np.int32)
        self._chk(lib().pb_scene_mesh_get_edges(self.h, mesh, _p(out)))
        return out

    # --- hot path ------------------------------------------------------------------------------------------------
    def trace(self, rays):
        """rays: (n, 8) float32 CUDA tensor (o.xyz, tmax, d.xyz, 0) -> (hits int32 (n,4) view pair, t)"""
        self._bind_stream()
        import torch
        n = rays.shape[0]
        hits = torch.empty((n, 4), dtype=torch.int32, device=rays.device)
        t = torch.empty(n, dtype=torch.float32, device=rays.device)
        self._chk(lib().pb_trace(self.h, C.c_int64(n), _dp(rays), _dp(hits), _dp(t)))
        return hits, t

    def trace_wavefront(self, rays):
        """The render calls' own ray launch (sort by direction / origin cell, compaction, streaming traversal kernel): rays (n, 8) float32
        CUDA tensor (o.xyz, tmax, d.xyz, t_occ) with origins inside the scene box -> hits int32 (n, 4)"""
        self._bind_stream()
        import torch
        n = rays.shape[0]
        hits = torch.empty((n, 4), dtype=torch.int32, device=rays.device)
        self._chk(lib().pb_trace(self.h, C.c_int64(n), _dp(rays), _dp(hits), None))
        return hits

    def primary_edges(self, sensor=0):
        """(n, 7) primary-edge table of a sensor as configure built it (p0.xy p1.xy normal.xy length) and its cmf"""
        n = self._id(lib().pb_scene_num_primary_edges(self.h, sensor))
        out, cmf = np.zeros((n, 7), np.float32), np.zeros(n, np.float32)
        self._chk(lib().pb_scene_get_primary_edges(self.h, sensor, _p(out), _p(cmf)))
        return out, cmf

    def secondary_edges(self):
        """(n, 16) secondary-edge table (p0 e1 n0 n1 p2 is_boundary) and its cmf"""
        n = self._id(lib().pb_scene_num_secondary_edges(self.h))
        out, cmf = np.zeros((n, 16), np.float32), np.zeros(n, np.float32)
        self._chk(lib().pb_scene_get_secondary_edges(self.h, _p(out), _p(cmf)))
        return out, cmf

    def sample_boundary_segment_direct(self, sample3):
        """Scene::sample_boundary_segment_direct: (n, 3) CUDA tensor of samples -> (n, 17): p0 edge edge2 p2 n pdf is_valid"""
        import torch
        self._bind_stream()
        s3 = sample3.contiguous().float()
        out = torch.empty((s3.shape[0], 17), dtype=torch.float32, device=s3.device)
        self._chk(lib().pb_sample_boundary_segment_direct(self.h, C.c_int64(s3.shape[0]), _dp(s3), _dp(out)))
        return out

    def _image(self):
        import torch
        return torch.empty((self.height * self.width, 3), dtype=torch.float32, device="cuda:%d" % self.device)

    def render_c(self, integ, sensor=0, out=None):
        self._bind_stream()
        img = out if out is not None else self._image()
        self._chk(lib().pb_render_c(self.h, C.byref(integ), sensor, _dp(img)))
        return img

    def render_c_host(self, integ, sensor=0, out=None):
        img = out if out is not None else np.empty((self.height * self.width, 3), dtype=np.float32)
        self._chk(lib().pb_render_c_host(self.h, C.byref(integ), sensor, _p(img)))
        return img

    def render_d(self, integ, sensor=0, out=None):
        self._bind_stream()
        img = out if out is not None else self._image()
        self._chk(lib().pb_render_d(self.h, C.byref(integ), sensor, _dp(img)))
        return img

    def preprocess_secondary_edges(self, sensor, resolution, nrounds=1):
        self._bind_stream()
        r = _i(resolution)
        assert r.shape == (4,)
        self._chk(lib().pb_preprocess_secondary_edges(self.h, sensor, _p(r), nrounds))

    # --- gradients ---------------------------------------------------------------------------------------------
    def grad_require(self, kind, id_, slot=0, enable=True):
        self._chk(lib().pb_grad_require(self.h, kind, id_, TEX[slot] if isinstance(slot, str) else slot, int(enable)))

    def grad_layout(self):
        L = lib()
        out = []
        for i in range(L.pb_grad_num_segments(self.h)):
            k, d, s = C.c_int(), C.c_int(), C.c_int()
            off, cnt = C.c_int64(), C.c_int64()
            self._chk(L.pb_grad_segment(self.h, i, C.byref(k), C.byref(d), C.byref(s), C.byref(off), C.byref(cnt)))
            out.append(dict(kind=k.value, id=d.value, slot=s.value, offset=off.value, count=cnt.value))
        return out

    def grad_size(self):
        return int(lib().pb_grad_size(self.h))

    def render_d_state(self):
        """replay state of the last render_d (sampler positions + serial number): pass it to render_d_vjp / render_d_jvp to
        differentiate that image after later renders (several sensors before one backward)"""
        st = (C.c_uint64 * 4)()
        self._chk(lib().pb_render_d_get_state(self.h, st))
        return tuple(int(x) for x in st)

    def _restore_state(self, state):
        if state is not None:
            st = (C.c_uint64 * 4)(*state)
            self._chk(lib().pb_render_d_set_state(self.h, st))

    def render_d_vjp(self, integ, dLdI, sensor=0, grad=None, state=None):
        import torch
        self._bind_stream()
        self._restore_state(state)
        if grad is None:
            grad = torch.zeros(max(1, self.grad_size()), dtype=torch.float32, device=dLdI.device)
        self._chk(lib().pb_render_d_vjp(self.h, C.byref(integ), sensor, _dp(dLdI.contiguous()), _dp(grad)))
        return grad

    def debug_set(self, key, value):
        self._chk(lib().pb_debug_set(self.h, key.encode(), C.c_int64(value)))

    def debug_ray_buffer(self, event):
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self._chk(lib().pb_debug_ray_buffer(self.h, event, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def render_d_jvp(self, integ, tangent, sensor=0, out=None, state=None):
        """forward mode: tangent is a flat CUDA tensor laid out like the gradient vector -> derivative image (W*H, 3)"""
        self._bind_stream()
        self._restore_state(state)
        img = out if out is not None else self._image()
        self._chk(lib().pb_render_d_jvp(self.h, C.byref(integ), sensor, _dp(tangent.contiguous()), _dp(img)))
        return img

    def set_bvh_refit(self, max_consecutive_refits):
        self._chk(lib().pb_ctx_set_bvh_refit(self.h, int(max_consecutive_refits)))

    def set_bvh_builder(self, builder):
        """0 = binned SAH on the host (default), 1 = LBVH on the device (pb_lbvh.cu)"""
        self._chk(lib().pb_ctx_set_bvh_builder(self.h, int(builder)))

    def bvh_stats(self):
        b, r = C.c_int(), C.c_int()
        self._chk(lib().pb_stats_bvh(self.h, C.byref(b), C.byref(r)))
        return dict(builds=b.value, refits=r.value)

    # --- stats -------------------------------------------------------------------------------------------------
    def stats(self):
        L = lib()
        return dict(launches=int(L.pb_stats_launches(self.h)), collectives=int(L.pb_stats_collectives(self.h)), trace_ms=float(L.pb_stats_last_trace_ms(self.h)), rays=int(L.pb_stats_last_rays(self.h)), active_rays=int(L.pb_stats_last_active_rays(self.h)),
                    trace_launches=int(L.pb_stats_last_trace_launches(self.h)), primary_ms=float(L.pb_stats_last_primary_ms(self.h)))
