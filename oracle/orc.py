"""ORACLE — TEST INFRASTRUCTURE ONLY (parity pinned against the reference's own source run on the CPU, oracle/_ref/libref_render.so: DESIGN.md §2).

ctypes driver + an independent Mitsuba-XML / OBJ loader for the CPU restatement in oracle/*.hpp.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this.

Loader follows (semantics, not code) src/scene/scene_loader.cpp:80-419 and src/shape/mesh.cpp:62-141 of
/root/reference; the OBJ triangulation rule is tinyobj's fan (a,b,c),(a,c,d) (pugixml/tinyobj row).
"""
import ctypes as C
import math
import os
import subprocess
import xml.etree.ElementTree as ET

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INTEG_DIRECT, INTEG_FIELD, INTEG_PATH = 0, 1, 2
FIELDS = {"silhouette": 0, "position": 1, "depth": 2, "geoNormal": 3, "shNormal": 4, "uv": 5}
BSDF_DIFFUSE, BSDF_ROUGHCONDUCTOR = 0, 1
TEX = {"reflectance": 0, "alpha_u": 1, "alpha_v": 2, "eta": 3, "k": 4, "specular_reflectance": 5}


def build(force=False):
    """Compile oracle/_build/liborc.so (gcc, -ffp-contract=off so the fp32 op order is the written one)."""
    out = os.path.join(_HERE, "_build", "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("orc_capi.cpp", "orc_math.hpp", "orc_scene.hpp", "orc_render.hpp")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fopenmp", "-shared", "-fPIC", "-ffp-contract=off",
                           "-o", out, srcs[0]])
    return out


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_last_error.restype = C.c_char_p
        L.orc_scene_new.restype = C.c_void_p
        L.orc_integrator_new.restype = C.c_void_p
        for name in ("orc_scene_free", "orc_integrator_free"):
            getattr(L, name).argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _chk(rc):
    if rc != 0:
        raise RuntimeError(lib().orc_last_error().decode())


# ------------------------------------------------------------------------------------------------
# Scene description loader (XML + OBJ) -> plain dict of numpy arrays
# ------------------------------------------------------------------------------------------------
def _parse_vec(s, n, allow_empty=False):
    vals = [float(x) for x in s.replace(",", " ").split()]
    if len(vals) < n:
        if not allow_empty:
            raise RuntimeError("Vector too short: [%s]" % s)
        vals = vals + [vals[-1] if vals else 0.0] * (n - len(vals))
    return np.array(vals[:n], dtype=np.float32)


def _f32(x):
    return np.float32(x)


def m_translate(v):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = v
    return m


def m_scale(v):
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[1, 1], m[2, 2] = v
    return m


def m_rotate(axis, angle_rad):
    a = np.asarray(axis, dtype=np.float32)
    s, c = _f32(math.sin(angle_rad)), _f32(math.cos(angle_rad))
    cm = _f32(1) - c
    m = np.eye(4, dtype=np.float32)
    x, y, z = a
    m[0, :3] = [x * x * cm + c, x * y * cm - z * s, x * z * cm + y * s]
    m[1, :3] = [y * x * cm + z * s, y * y * cm + c, y * z * cm - x * s]
    m[2, :3] = [z * x * cm - y * s, z * y * cm + x * s, z * z * cm + c]
    return m


def m_look_at(origin, target, up):
    def nrm(v):
        return (v / np.sqrt(np.dot(v, v), dtype=np.float32)).astype(np.float32)
    d = nrm(target - origin)
    left = nrm(np.cross(up, d).astype(np.float32))
    new_up = np.cross(d, left).astype(np.float32)
    m = np.eye(4, dtype=np.float32)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, origin
    return m


def _load_transform(node):
    result = np.eye(4, dtype=np.float32)
    if node is None:
        return result
    if node.get("name") not in ("to_world", "toWorld"):
        raise RuntimeError("Invalid transformation name: %s" % node.get("name"))
    for ch in node:
        g = lambda k, d: float(ch.get(k, d))
        if ch.tag == "translate":
            t = m_translate([g("x", 0), g("y", 0), g("z", 0)])
        elif ch.tag == "rotate":
            t = m_rotate([g("x", 0), g("y", 0), g("z", 0)], float(np.float32(g("angle", 0)) * (np.float32(math.pi) / np.float32(180))))
        elif ch.tag == "scale":
            t = m_scale([g("x", 1), g("y", 1), g("z", 1)])
        elif ch.tag in ("look_at", "lookAt", "lookat"):
            t = m_look_at(_parse_vec(ch.get("origin"), 3), _parse_vec(ch.get("target"), 3), _parse_vec(ch.get("up"), 3))
        elif ch.tag == "matrix":
            t = _parse_vec(ch.get("value"), 16).reshape(4, 4)
        else:
            raise RuntimeError("Unsupported transformation: %s" % ch.tag)
        result = (t @ result).astype(np.float32)
    return result


def _find_named(node, names):
    for ch in node:
        if ch.get("name") in names:
            return ch
    return None


def resolve_path(fname, xml_dir):
    cands = [fname]
    d = xml_dir
    for _ in range(4):
        if d:
            cands.append(os.path.join(d, fname))
            d = os.path.dirname(d)
    for c in cands:
        if os.path.exists(c):
            return c
    raise RuntimeError("Failed to load: %s" % fname)


def load_obj(path):
    """tinyobj-style ingest: positions, optional texcoords, fan triangulation, per-corner (v, vt) indices."""
    vs, vts, fv, ft = [], [], [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                vs.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("vt "):
                p = line.split()
                vts.append((float(p[1]), float(p[2]) if len(p) > 2 else 0.0))
            elif line.startswith("f "):
                corners = []
                for tok in line.split()[1:]:
                    parts = tok.split("/")
                    vi = int(parts[0])
                    vi = vi - 1 if vi > 0 else len(vs) + vi
                    ti = -1
                    if len(parts) > 1 and parts[1]:
                        ti = int(parts[1])
                        ti = ti - 1 if ti > 0 else len(vts) + ti
                    corners.append((vi, ti))
                for k in range(1, len(corners) - 1):
                    tri = (corners[0], corners[k], corners[k + 1])
                    fv.append([c[0] for c in tri])
                    ft.append([c[1] for c in tri])
    d = {"verts": np.array(vs, dtype=np.float32).reshape(-1, 3), "faces": np.array(fv, dtype=np.int32).reshape(-1, 3)}
    if vts:
        d["uvs"] = np.array(vts, dtype=np.float32).reshape(-1, 2)
        d["uv_faces"] = np.array(ft, dtype=np.int32).reshape(-1, 3)
    return d


def load_exr(path):
    import cv2
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise RuntimeError("Failed to load EXR: %s" % path)
    return np.ascontiguousarray(img[:, :, 2::-1].astype(np.float32)) if img.ndim == 3 else img.astype(np.float32)[:, :, None]


def _texture(node, channels, xml_dir):
    """-> (h, w, channels) float32 array (1x1 for constants). scene_loader.cpp:147-172"""
    if node.tag == "texture":
        fn = node.find("string").get("value")
        img = load_exr(resolve_path(fn, xml_dir))
        return np.ascontiguousarray(img[:, :, :channels])
    if channels == 1:
        return np.full((1, 1, 1), float(node.get("value")), dtype=np.float32)
    if node.tag == "float":
        return np.full((1, 1, 3), float(node.get("value")), dtype=np.float32)
    if node.tag == "rgb":
        return _parse_vec(node.get("value"), 3, True).reshape(1, 1, 3)
    raise RuntimeError("Unsupported RGB type: %s" % node.tag)


def load_scene_description(xml_path=None, xml_string=None):
    if xml_path is not None:
        root = ET.parse(xml_path).getroot()
        xml_dir = os.path.dirname(os.path.abspath(xml_path))
    else:
        root = ET.fromstring(xml_string)
        xml_dir = os.getcwd()
    desc = {"sensors": [], "bsdfs": [], "meshes": [], "emitters": [], "envmap": None, "opts": None}
    for node in root.findall("sensor"):
        film, sampler = node.find("film"), node.find("sampler")
        if not desc["sensors"]:            # scene_loader.cpp:251-262: film and sampler belong to the first sensor, and only to it
            if film is None:
                raise RuntimeError("Missing film node")
            if sampler is None:
                raise RuntimeError("Missing sampler node")
            w = int(_find_named(film, ("width",)).get("value"))
            h = int(_find_named(film, ("height",)).get("value"))
            spp = int(sampler.find("integer").get("value"))
            desc["opts"] = dict(width=w, height=h, spp=spp, sppe=spp, sppse=spp)
        else:
            if film is not None:
                raise RuntimeError("Duplicate film node")
            if sampler is not None:
                raise RuntimeError("Duplicate sampler node")
        if node.get("type") != "perspective":
            raise RuntimeError("Unsupported sensor: %s" % node.get("type"))
        fa = _find_named(node, ("fov_axis", "fovAxis"))
        if fa is not None and fa.get("value") != "x":
            raise RuntimeError("Unsupported fov-axis: %s" % fa.get("value"))
        nn, fn_ = _find_named(node, ("near_clip", "nearClip")), _find_named(node, ("far_clip", "farClip"))
        desc["sensors"].append(dict(fov=float(_find_named(node, ("fov",)).get("value")),
                                    near=float(nn.get("value")) if nn is not None else 0.1,
                                    far=float(fn_.get("value")) if fn_ is not None else 1e4,
                                    to_world=_load_transform(node.find("transform"))))
    bsdf_ids = {}
    for node in root.findall("bsdf"):
        bid = node.get("id")
        if not bid:
            raise RuntimeError("BSDF must have an id")
        if node.get("type") == "diffuse":
            b = dict(type=BSDF_DIFFUSE, id=bid, reflectance=_texture(_find_named(node, ("reflectance",)), 3, xml_dir))
        elif node.get("type") == "roughconductor":
            alpha = _texture(_find_named(node, ("alpha",)), 1, xml_dir)
            b = dict(type=BSDF_ROUGHCONDUCTOR, id=bid, alpha_u=alpha, alpha_v=alpha.copy(),
                     eta=_texture(_find_named(node, ("eta",)), 3, xml_dir), k=_texture(_find_named(node, ("k",)), 3, xml_dir))
        else:
            raise RuntimeError("Unsupported BSDF: %s" % node.get("type"))
        if bid in bsdf_ids:
            raise RuntimeError("Duplicate BSDF id: %s" % bid)
        bsdf_ids[bid] = len(desc["bsdfs"])
        desc["bsdfs"].append(b)
    for node in root.findall("emitter"):
        if node.get("type") != "envmap":
            raise RuntimeError("Unsupported emitter: %s" % node.get("type"))
        sc = _find_named(node, ("scale",))
        desc["envmap"] = dict(radiance=load_exr(resolve_path(node.find("string").get("value"), xml_dir))[:, :, :3].copy(),
                              scale=float(sc.get("value")) if sc is not None else 1.0,
                              to_world=_load_transform(node.find("transform")))
    for node in root.findall("shape"):
        if node.get("type") != "obj":
            raise RuntimeError("Unsupported shape: %s" % node.get("type"))
        m = load_obj(resolve_path(node.find("string").get("value"), xml_dir))
        ref = node.find("ref")
        if ref is None:
            raise RuntimeError("Missing BSDF reference")
        if ref.get("id") not in bsdf_ids:
            raise RuntimeError("Unknown BSDF id: %s" % ref.get("id"))
        m["bsdf"] = bsdf_ids[ref.get("id")]
        fnn = _find_named(node, ("face_normals", "faceNormals"))
        m["face_normals"] = fnn is not None and fnn.get("value") == "true"
        m["enable_edges"] = True
        m["id"] = node.get("id") or ""
        m["to_world"] = _load_transform(node.find("transform"))
        em = node.find("emitter")
        if em is not None:
            if em.get("type") != "area":
                raise RuntimeError("Unsupported emitter: %s" % em.get("type"))
            rad = _texture(_find_named(em, ("radiance",)), 3, xml_dir).reshape(3)
            desc["emitters"].append(dict(mesh=len(desc["meshes"]), radiance=rad))
        desc["meshes"].append(m)
    return desc


# ------------------------------------------------------------------------------------------------
# Oracle objects
# ------------------------------------------------------------------------------------------------
class Scene:
    def __init__(self, desc, opts=None):
        L = lib()
        self.h = C.c_void_p(L.orc_scene_new())
        self.desc = desc
        self.opts = dict(desc["opts"])
        if opts:
            self.opts.update(opts)
        # the reference order: sensors, bsdfs, env emitter, shapes (+ their area emitters)
        for s in desc["sensors"]:
            L.orc_add_sensor(self.h, C.c_float(s["fov"]), C.c_float(s["near"]), C.c_float(s["far"]), _p(_f(s["to_world"])))
        for b in desc["bsdfs"]:
            bi = L.orc_add_bsdf(self.h, b["type"])
            for k, which in TEX.items():
                if k in b:
                    t = _f(b[k])
                    _chk(L.orc_set_bsdf_texture(self.h, bi, which, _p(t), t.shape[1], t.shape[0]))
        if desc["envmap"] is not None:
            e = desc["envmap"]
            r = _f(e["radiance"])
            L.orc_add_envmap(self.h, r.shape[1], r.shape[0], _p(r), C.c_float(e["scale"]), _p(_f(e["to_world"])))
        emitter_of = {e["mesh"]: e for e in desc["emitters"]}
        for mi, m in enumerate(desc["meshes"]):
            has_uv = "uvs" in m
            rc = L.orc_add_mesh(self.h, len(m["verts"]), len(m["faces"]), _p(_f(m["verts"])), _p(_i(m["faces"])),
                                len(m["uvs"]) if has_uv else 0, _p(_f(m["uvs"])) if has_uv else None,
                                _p(_i(m["uv_faces"])) if has_uv else None, int(m["face_normals"]), int(m["enable_edges"]),
                                m["bsdf"], _p(_f(m["to_world"])))
            if rc < 0:
                _chk(1)
            if mi in emitter_of:
                L.orc_add_area_emitter(self.h, mi, _p(_f(emitter_of[mi]["radiance"])))

    def __del__(self):
        try:
            lib().orc_scene_free(self.h)
        except Exception:
            pass

    def configure(self, reseed=False):
        L = lib()
        if reseed:
            L.orc_reseed(self.h)
        o = self.opts
        L.orc_set_options(self.h, o["width"], o["height"], o["spp"], o["sppe"], o["sppse"])
        _chk(L.orc_configure(self.h))

    # --- parameters and tangents -------------------------------------------------------------------
    def set_bsdf_texture(self, bsdf, name, data):
        t = _f(data)
        _chk(lib().orc_set_bsdf_texture(self.h, bsdf, TEX[name], _p(t), t.shape[1], t.shape[0]))

    def set_bsdf_tangent(self, bsdf, name, tang):
        _chk(lib().orc_set_bsdf_tangent(self.h, bsdf, TEX[name], _p(_f(tang)) if tang is not None else None))

    def set_mesh_uv_tangent(self, mesh, tang):
        L = lib()
        L.orc_set_mesh_uv_tangent.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _chk(L.orc_set_mesh_uv_tangent(self.h, mesh, _p(_f(tang)) if tang is not None else None))

    def set_sensor_transform_tangent(self, sensor, tang):
        L = lib()
        L.orc_set_sensor_transform_tangent.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _chk(L.orc_set_sensor_transform_tangent(self.h, sensor, _p(_f(tang)) if tang is not None else None))

    def set_envmap_transform_tangent(self, tang):
        L = lib()
        L.orc_set_envmap_transform_tangent.argtypes = [C.c_void_p, C.c_void_p]
        _chk(L.orc_set_envmap_transform_tangent(self.h, _p(_f(tang)) if tang is not None else None))

    def set_envmap_tangent(self, radiance_t=None, scale_t=0.0):
        L = lib()
        L.orc_set_envmap_tangent.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        _chk(L.orc_set_envmap_tangent(self.h, _p(_f(radiance_t)) if radiance_t is not None else None, float(scale_t)))

    def set_mesh_vertices(self, mesh, verts):
        _chk(lib().orc_set_mesh_vertices(self.h, mesh, _p(_f(verts))))

    def set_mesh_vertex_tangent(self, mesh, tang):
        _chk(lib().orc_set_mesh_vertex_tangent(self.h, mesh, _p(_f(tang)) if tang is not None else None))

    def set_mesh_transform(self, mesh, mat, left=True):
        _chk(lib().orc_set_mesh_transform(self.h, mesh, _p(_f(mat)), int(left)))

    def set_mesh_transform_tangent(self, mesh, tang, left=True):
        _chk(lib().orc_set_mesh_transform_tangent(self.h, mesh, _p(_f(tang)) if tang is not None else None, int(left)))

    # --- introspection --------------------------------------------------------------------------------
    def triangle_info(self):
        n = lib().orc_num_triangles(self.h)
        out = np.empty((n, 22), dtype=np.float32)
        lib().orc_get_triangle_info(self.h, _p(out))
        return out

    def mesh_edges(self, mesh):
        n = lib().orc_mesh_num_edges(self.h, mesh)
        out = np.empty((n, 5), dtype=np.int32)
        lib().orc_mesh_get_edges(self.h, mesh, _p(out))
        return out

    def sec_edges(self):
        n = lib().orc_num_sec_edges(self.h)
        out = np.empty((n, 16), dtype=np.float32)
        lib().orc_get_sec_edges(self.h, _p(out))
        return out

    def primary_edges(self, sensor=0):
        n = lib().orc_num_primary_edges(self.h, sensor)
        out = np.empty((n, 7), dtype=np.float32)
        lib().orc_get_primary_edges(self.h, sensor, _p(out))
        return out

    def sensor_info(self, sensor=0):
        out = np.empty(55, dtype=np.float32)
        lib().orc_get_sensor(self.h, sensor, _p(out))
        return dict(sample_to_camera=out[:16].reshape(4, 4), world_to_sample=out[16:32].reshape(4, 4),
                    to_world=out[32:48].reshape(4, 4), camera_pos=out[48:51], camera_dir=out[51:54], inv_area=out[54])

    def trace(self, o, d, tmax=None, brute=False):
        o, d = _f(o), _f(d)
        n = len(o)
        tri, shape = np.empty(n, np.int32), np.empty(n, np.int32)
        u, v, t = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
        tm = _f(tmax) if tmax is not None else None
        _chk(lib().orc_trace(self.h, C.c_int64(n), _p(o), _p(d), _p(tm), _p(tri), _p(shape), _p(u), _p(v), _p(t), int(brute)))
        return tri, shape, u, v, t


class Integrator:
    def __init__(self, kind=INTEG_DIRECT, bsdf_samples=1, light_samples=1, hide_emitters=False, field="silhouette", max_depth=1):
        self.h = C.c_void_p(lib().orc_integrator_new(kind, bsdf_samples, light_samples, int(hide_emitters), FIELDS[field], max_depth))

    def __del__(self):
        try:
            lib().orc_integrator_free(self.h)
        except Exception:
            pass

    def renderC(self, scene, sensor=0):
        o = scene.opts
        out = np.zeros((o["height"] * o["width"], 3), dtype=np.float32)
        _chk(lib().orc_render_c(scene.h, self.h, sensor, _p(out)))
        return out

    def renderD(self, scene, sensor=0):
        """-> (image, tangent image) for the tangents currently set on the scene (forward mode)."""
        o = scene.opts
        out = np.zeros((o["height"] * o["width"], 3), dtype=np.float32)
        out_t = np.zeros_like(out)
        _chk(lib().orc_render_d(scene.h, self.h, sensor, _p(out), _p(out_t)))
        return out, out_t

    def preprocess_secondary_edges(self, scene, sensor, reso, nrounds=1):
        r = _i(reso)
        _chk(lib().orc_preprocess_secondary_edges(scene.h, self.h, sensor, _p(r), nrounds))


def DirectIntegrator(bsdf_samples=1, light_samples=1, hide_emitters=False):
    return Integrator(INTEG_DIRECT, bsdf_samples, light_samples, hide_emitters)


def FieldExtractionIntegrator(field):
    return Integrator(INTEG_FIELD, field=field)


def PathIntegrator(max_depth=1, hide_emitters=False):
    return Integrator(INTEG_PATH, hide_emitters=hide_emitters, max_depth=max_depth)


def pcg32_kat(initstate, initseq, n):
    out = np.empty(n, np.uint32)
    lib().orc_pcg32_kat(C.c_uint64(initstate), C.c_uint64(initseq), n, _p(out))
    return out


def sampler_kat(lane, n):
    u, f = np.empty(n, np.uint32), np.empty(n, np.float32)
    lib().orc_sampler_kat(C.c_uint64(lane), n, _p(u), _p(f))
    return u, f
