"""TEST INFRASTRUCTURE: ctypes driver for oracle/_ref/libref_render.so = psdr-cuda's OWN renderer sources (Scene, SceneLoader, Mesh,
PerspectiveCamera, Diffuse / RoughConductor / GGX, AreaLight / EnvironmentMap, DiscreteDistribution / HyperCubeDistribution, Sampler,
Integrator / DirectIntegrator / FieldExtractionIntegrator) compiled unmodified from /root/reference against a CPU stand-in for Enoki
(oracle/ref_dyn/enoki_dyn.h) and for the OptiX glue (oracle/ref_render_shim.cpp). Built by oracle/build_ref.sh.

Only tests/ may import this. It mirrors the call shapes of oracle/orc.py (Scene / Integrator, tangent setters, table getters) so that a
test can run the same steps through the reference's code and through the oracle and compare."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_render.so")
REFERENCE = os.environ.get("PSDR_REFERENCE", "/root/reference")
TEX = {"reflectance": 0, "alpha_u": 1, "alpha_v": 2, "eta": 3, "k": 4, "specular_reflectance": 5}
_LIB = None


def available():
    """the library exists, or can be built because the reference sources are here (this container; never on the GPU box)"""
    return os.path.exists(LIB_PATH) or os.path.isdir(os.path.join(REFERENCE, "src"))


def lib():
    global _LIB
    if _LIB is None:
        srcs = [os.path.join(_HERE, f) for f in ("ref_render_shim.cpp", "ref_dyn/enoki_dyn.h", "build_ref_render.sh")]
        stale = os.path.exists(LIB_PATH) and any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
        if (stale or not os.path.exists(LIB_PATH)) and os.path.isdir(os.path.join(REFERENCE, "src")):
            subprocess.check_call(["bash", os.path.join(_HERE, "build_ref_render.sh")], stdout=subprocess.DEVNULL)
        L = C.CDLL(LIB_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_scene_load.restype = C.c_void_p
        L.ref_scene_load.argtypes = [C.c_char_p, C.c_char_p] + [C.c_int] * 5
        L.ref_integrator_new.restype = C.c_void_p
        L.ref_integrator_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
        for name in ("ref_scene_free", "ref_integrator_free"):
            getattr(L, name).argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _chk(rc):
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())


def set_matvec_plain(on):
    """matrix * vector as plain sums (the oracle's form) instead of the fmadd chain assumed for Enoki: see oracle/ref_dyn/enoki_dyn.h"""
    lib().ref_set_matvec_plain(int(on))


def set_inverse_rounded(on):
    """4x4 inverses computed in double precision and rounded (the oracle's / the product host's form) instead of an fp32 cofactor expansion"""
    lib().ref_set_inverse_rounded(int(on))


def set_dot_from_first(on):
    """dot products accumulated from the first component up instead of from the last down (see oracle/ref_dyn/enoki_dyn.h)"""
    lib().ref_set_dot_from_first(int(on))


class Scene:
    """psdr::Scene: load_file(xml, auto_configure=False) + RenderOption overrides; configure() is explicit, as in the reference"""

    def __init__(self, xml_path, cwd, width=0, height=0, spp=-1, sppe=0, sppse=0):
        L = lib()
        h = L.ref_scene_load(os.path.abspath(xml_path).encode(), os.path.abspath(cwd).encode(), width, height, spp, sppe, sppse)
        if not h:
            raise RuntimeError(L.ref_last_error().decode())
        self.h = C.c_void_p(h)
        o = np.zeros(5, np.int32)
        L.ref_scene_options(self.h, _p(o))
        self.opts = dict(width=int(o[0]), height=int(o[1]), spp=int(o[2]), sppe=int(o[3]), sppse=int(o[4]))

    def __del__(self):
        try:
            lib().ref_scene_free(self.h)
        except Exception:
            pass

    def configure(self):
        _chk(lib().ref_scene_configure(self.h))

    # tangents (forward mode) and parameter values
    def set_bsdf_tangent(self, bsdf, name, tang):
        _chk(lib().ref_set_bsdf_tangent(self.h, bsdf, TEX[name], _p(_f(tang))))

    def set_bsdf_texture(self, bsdf, name, data):
        t = _f(data)
        _chk(lib().ref_set_bsdf_texture(self.h, bsdf, TEX[name], _p(t), t.shape[1], t.shape[0]))

    def num_vertices(self, mesh):
        return lib().ref_mesh_num_vertices(self.h, mesh)

    def set_mesh_vertex_tangent(self, mesh, tang):
        _chk(lib().ref_set_mesh_vertex_tangent(self.h, mesh, _p(_f(tang))))

    def set_mesh_uv_tangent(self, mesh, tang):
        _chk(lib().ref_set_mesh_uv_tangent(self.h, mesh, _p(_f(tang))))

    def set_mesh_transform(self, mesh, mat, left=True):
        _chk(lib().ref_set_mesh_transform(self.h, mesh, _p(_f(mat)), int(left)))

    def set_mesh_transform_tangent(self, mesh, tang, left=True):
        _chk(lib().ref_set_mesh_transform_tangent(self.h, mesh, _p(_f(tang)), int(left)))

    def set_sensor_transform_tangent(self, sensor, tang):
        _chk(lib().ref_set_sensor_transform_tangent(self.h, sensor, _p(_f(tang))))

    def set_envmap_tangent(self, radiance_t=None, scale_t=0.0):
        _chk(lib().ref_set_envmap_tangent(self.h, _p(_f(radiance_t)), C.c_float(scale_t)))

    def set_envmap_transform_tangent(self, tang):
        _chk(lib().ref_set_envmap_transform_tangent(self.h, _p(_f(tang))))

    # tables, in the layouts of oracle/orc.py
    def triangle_info(self):
        out = np.zeros((lib().ref_num_triangles(self.h), 22), np.float32)
        lib().ref_get_triangle_info(self.h, _p(out))
        return out

    def sec_edges(self):
        out = np.zeros((lib().ref_num_sec_edges(self.h), 16), np.float32)
        if len(out):
            lib().ref_get_sec_edges(self.h, _p(out))
        return out

    def primary_edges(self, sensor=0):
        out = np.zeros((lib().ref_num_primary_edges(self.h, sensor), 7), np.float32)
        if len(out):
            lib().ref_get_primary_edges(self.h, sensor, _p(out))
        return out

    def mesh_edges(self, mesh):
        out = np.zeros((lib().ref_mesh_num_edges(self.h, mesh), 5), np.int32)
        if len(out):
            lib().ref_mesh_get_edges(self.h, mesh, _p(out))
        return out

    def num_meshes(self):
        return lib().ref_scene_num_meshes(self.h)

    def sensor_info(self, sensor=0):
        out = np.zeros(55, np.float32)
        _chk(lib().ref_get_sensor(self.h, sensor, _p(out)))
        return dict(sample_to_camera=out[:16].reshape(4, 4), world_to_sample=out[16:32].reshape(4, 4), to_world=out[32:48].reshape(4, 4),
                    camera_pos=out[48:51], camera_dir=out[51:54], inv_area=out[54])

    def sample_boundary_segment_direct(self, sample3):
        """Scene::sample_boundary_segment_direct (scene.cpp:456-492) -> (n, 17): p0 edge edge2 p2 n pdf is_valid"""
        s3 = _f(sample3)
        out = np.zeros((len(s3), 17), np.float32)
        _chk(lib().ref_sample_boundary_segment_direct(self.h, C.c_int64(len(s3)), _p(s3), _p(out)))
        return out

    def trace(self, o, d):
        o, d = _f(o), _f(d)
        n = len(o)
        tri, shape = np.zeros(n, np.int32), np.zeros(n, np.int32)
        u, v, t = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        _chk(lib().ref_trace(self.h, C.c_int64(n), _p(o), _p(d), _p(tri), _p(shape), _p(u), _p(v), _p(t)))
        return tri, shape, u, v, t


class Integrator:
    def __init__(self, kind=0, bsdf_samples=1, light_samples=1, hide_emitters=False, field="silhouette"):
        h = lib().ref_integrator_new(kind, bsdf_samples, light_samples, int(hide_emitters), field.encode())
        if not h:
            raise RuntimeError(lib().ref_last_error().decode())
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            lib().ref_integrator_free(self.h)
        except Exception:
            pass

    def renderC(self, scene, sensor=0):
        out = np.zeros((scene.opts["width"] * scene.opts["height"], 3), np.float32)
        _chk(lib().ref_render_c(scene.h, self.h, sensor, _p(out)))
        return out

    def renderD(self, scene, sensor=0):
        """-> (image, forward-mode tangent image) for the tangents set on the scene"""
        out = np.zeros((scene.opts["width"] * scene.opts["height"], 3), np.float32)
        out_t = np.zeros_like(out)
        _chk(lib().ref_render_d(scene.h, self.h, sensor, _p(out), _p(out_t)))
        return out, out_t

    def lane_radiance(self, scene, sensor=0, ad=False, skip_dims=0):
        """per-lane radiance of one interior pass, before the scatter_add to pixels (advances sampler 0 like a render)"""
        o = scene.opts
        out = np.zeros((o["width"] * o["height"] * o["spp"], 3), np.float32)
        _chk(lib().ref_debug_li(scene.h, self.h, sensor, int(ad), skip_dims, _p(out)))
        return out

    def preprocess_secondary_edges(self, scene, sensor, reso, nrounds=1):
        r = np.ascontiguousarray(reso, dtype=np.int32)
        _chk(lib().ref_preprocess_secondary_edges(scene.h, self.h, sensor, _p(r), nrounds))


def DirectIntegrator(bsdf_samples=1, light_samples=1, hide_emitters=False):
    return Integrator(0, bsdf_samples, light_samples, hide_emitters)


def FieldExtractionIntegrator(field):
    return Integrator(1, field=field)
