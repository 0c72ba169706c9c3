"""Scene ingest for the product path: Mitsuba-style XML subset + Wavefront OBJ -> scene-description dict.

Mirrors what SceneLoader::load_scene (src/scene/scene_loader.cpp:208-419) and Mesh::load (src/shape/mesh.cpp:62-141)
accept in the reference: sensor `perspective` (fov axis x only), bsdf `diffuse` / `roughconductor`, shape `obj`
with an optional `area` emitter, top-level `envmap` emitter, transforms translate / rotate (degrees) / scale /
lookat / matrix composed by left-multiplication in document order. Errors are RuntimeError with the reference's
messages. File names resolve against the working directory first (as the reference does) and then against the
directories above the XML file.
"""
import math
import os
import xml.dom.minidom as minidom

import numpy as np

F32 = np.float32


class SceneFormatError(RuntimeError):
    pass


def _floats(text, n, pad=False):
    parts = [p for p in text.replace(",", " ").split(" ") if p]
    vals = [float(p) for p in parts]
    if len(vals) > n:
        raise SceneFormatError("Vector too long: [%s]" % text)
    if len(vals) < n:
        if not pad:
            raise SceneFormatError("Vector too short: [%s]" % text)
        fill = vals[-1] if vals else 0.0
        vals += [fill] * (n - len(vals))
    return np.asarray(vals, dtype=F32)


def _elements(node, tag=None):
    return [c for c in node.childNodes if c.nodeType == c.ELEMENT_NODE and (tag is None or c.tagName == tag)]


def _named(node, *names):
    for c in _elements(node):
        if c.getAttribute("name") in names:
            return c
    return None


def _attr_float(node, key, default):
    return float(node.getAttribute(key)) if node.hasAttribute(key) else default


class Transform:
    """4x4 row-major fp32 matrices with the reference's constructors (include/psdr/core/transform.h)."""

    @staticmethod
    def translate(x, y, z):
        m = np.identity(4, dtype=F32)
        m[0, 3], m[1, 3], m[2, 3] = x, y, z
        return m

    @staticmethod
    def scale(x, y, z):
        return np.diag(np.asarray([x, y, z, 1], dtype=F32))

    @staticmethod
    def rotate(axis, angle_deg):
        ang = float(F32(angle_deg) * (F32(math.pi) / F32(180)))   # transform.h:27: deg_to_rad(a) = a * (Pi / 180)
        s, c = F32(math.sin(ang)), F32(math.cos(ang))
        x, y, z = (F32(v) for v in axis)
        k = F32(1) - c
        m = np.identity(4, dtype=F32)
        m[:3, :3] = [[x * x * k + c, x * y * k - z * s, x * z * k + y * s],
                     [y * x * k + z * s, y * y * k + c, y * z * k - x * s],
                     [z * x * k - y * s, z * y * k + x * s, z * z * k + c]]
        return m

    @staticmethod
    def look_at(origin, target, up):
        def unit(v):
            v = v.astype(F32)
            return v / np.sqrt((v * v).sum(dtype=F32), dtype=F32)
        fwd = unit(target - origin)
        left = unit(np.cross(up, fwd))
        new_up = np.cross(fwd, left).astype(F32)
        m = np.identity(4, dtype=F32)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, fwd, origin
        return m

    @staticmethod
    def from_xml(node):
        m = np.identity(4, dtype=F32)
        if node is None:
            return m
        if node.getAttribute("name") not in ("to_world", "toWorld"):
            raise SceneFormatError("Invalid transformation name: " + node.getAttribute("name"))
        for op in _elements(node):
            tag = op.tagName
            if tag == "translate":
                t = Transform.translate(_attr_float(op, "x", 0.0), _attr_float(op, "y", 0.0), _attr_float(op, "z", 0.0))
            elif tag == "rotate":
                t = Transform.rotate((_attr_float(op, "x", 0.0), _attr_float(op, "y", 0.0), _attr_float(op, "z", 0.0)), _attr_float(op, "angle", 0.0))
            elif tag == "scale":
                t = Transform.scale(_attr_float(op, "x", 1.0), _attr_float(op, "y", 1.0), _attr_float(op, "z", 1.0))
            elif tag in ("look_at", "lookAt", "lookat"):
                t = Transform.look_at(_floats(op.getAttribute("origin"), 3), _floats(op.getAttribute("target"), 3), _floats(op.getAttribute("up"), 3))
            elif tag == "matrix":
                t = _floats(op.getAttribute("value"), 16).reshape(4, 4)
            else:
                raise SceneFormatError("Unsupported transformation: " + tag)
            m = np.matmul(t, m, dtype=F32)
        return m


def find_file(name, xml_dir):
    trial = [name]
    d = xml_dir
    while d and len(trial) < 6:
        trial.append(os.path.join(d, name))
        parent = os.path.dirname(d)
        if parent == d:
            break
        d = parent
    for t in trial:
        if os.path.isfile(t):
            return t
    raise SceneFormatError("Failed to load file: " + name)


def read_obj(path):
    """positions, optional texcoords, polygons fanned as (a,b,c),(a,c,d) like tinyobj's triangulation of convex faces"""
    pos, tex, tri_v, tri_t = [], [], [], []
    with open(path, "r") as fh:
        for raw in fh:
            if len(raw) < 2:
                continue
            head = raw[:2]
            if head == "v ":
                pos.append(raw.split()[1:4])
            elif head == "vt":
                tex.append((raw.split()[1:3] + ["0"])[:2])
            elif head == "f ":
                vi, ti = [], []
                for corner in raw.split()[1:]:
                    ref = corner.split("/")
                    a = int(ref[0])
                    vi.append(a - 1 if a > 0 else len(pos) + a)
                    if len(ref) > 1 and ref[1] != "":
                        b = int(ref[1])
                        ti.append(b - 1 if b > 0 else len(tex) + b)
                    else:
                        ti.append(-1)
                for k in range(2, len(vi)):
                    tri_v.append((vi[0], vi[k - 1], vi[k]))
                    tri_t.append((ti[0], ti[k - 1], ti[k]))
    if not pos:
        raise SceneFormatError("Failed to load OBJ from: " + path)
    mesh = {"verts": np.asarray(pos, dtype=F32).reshape(-1, 3), "faces": np.asarray(tri_v, dtype=np.int32).reshape(-1, 3)}
    if tex:
        mesh["uvs"] = np.asarray(tex, dtype=F32).reshape(-1, 2)
        mesh["uv_faces"] = np.asarray(tri_t, dtype=np.int32).reshape(-1, 3)
    return mesh


def read_exr(path):
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise SceneFormatError("Failed to load EXR: " + path)
    img = img.astype(F32)
    if img.ndim == 2:
        img = img[:, :, None]
    else:
        img = img[:, :, [2, 1, 0] + list(range(3, img.shape[2]))]
    return np.ascontiguousarray(img)


def _texture_value(node, channels, xml_dir):
    if node.tagName == "texture":
        if node.getAttribute("type") != "bitmap":
            raise SceneFormatError("Unsupported texture type: " + node.getAttribute("type"))
        fn = _elements(node, "string")[0]
        if fn.getAttribute("name") != "filename":
            raise SceneFormatError("Failed to retrieve bitmap filename")
        return np.ascontiguousarray(read_exr(find_file(fn.getAttribute("value"), xml_dir))[:, :, :channels])
    if channels == 1:
        return np.full((1, 1, 1), float(node.getAttribute("value")), dtype=F32)
    if node.tagName == "float":
        return np.full((1, 1, 3), float(node.getAttribute("value")), dtype=F32)
    if node.tagName == "rgb":
        return _floats(node.getAttribute("value"), 3, pad=True).reshape(1, 1, 3)
    raise SceneFormatError("Unsupported RGB type: " + node.tagName)


def _require(node, what, *names):
    r = _named(node, *names)
    if r is None:
        raise SceneFormatError("Missing child node: " + what)
    return r


def load_scene_description(xml_path=None, xml_string=None):
    if xml_path is not None:
        try:
            dom = minidom.parse(xml_path)
        except Exception:
            raise SceneFormatError("XML parsing failed")
        xml_dir = os.path.dirname(os.path.abspath(xml_path))
    else:
        try:
            dom = minidom.parseString(xml_string)
        except Exception:
            raise SceneFormatError("XML parsing failed")
        xml_dir = os.getcwd()
    roots = [n for n in dom.childNodes if n.nodeType == n.ELEMENT_NODE and n.tagName == "scene"]
    if not roots:
        raise SceneFormatError("XML parsing failed")
    root = roots[0]
    scene = {"opts": None, "sensors": [], "bsdfs": [], "meshes": [], "emitters": [], "envmap": None}

    for node in _elements(root, "sensor"):
        films, samplers = _elements(node, "film"), _elements(node, "sampler")
        if not scene["sensors"]:
            if not films:
                raise SceneFormatError("Missing film node")
            if not samplers:
                raise SceneFormatError("Missing sampler node")
            width = int(_require(films[0], "width", "width").getAttribute("value"))
            height = int(_require(films[0], "height", "height").getAttribute("value"))
            count = int(_elements(samplers[0], "integer")[0].getAttribute("value"))
            scene["opts"] = {"width": width, "height": height, "spp": count, "sppe": count, "sppse": count}
        else:
            if films:
                raise SceneFormatError("Duplicate film node")
            if samplers:
                raise SceneFormatError("Duplicate sampler node")
        if node.getAttribute("type") != "perspective":
            raise SceneFormatError("Unsupported sensor: " + node.getAttribute("type"))
        axis = _named(node, "fov_axis", "fovAxis")
        if axis is not None and axis.getAttribute("value") != "x":
            raise SceneFormatError("Unsupported fov-axis: " + axis.getAttribute("value"))
        near, far = _named(node, "near_clip", "nearClip"), _named(node, "far_clip", "farClip")
        tr = _elements(node, "transform")
        scene["sensors"].append({"fov": float(_require(node, "fov", "fov").getAttribute("value")),
                                 "near": float(near.getAttribute("value")) if near is not None else 0.1,
                                 "far": float(far.getAttribute("value")) if far is not None else 1e4,
                                 "to_world": Transform.from_xml(tr[0] if tr else None)})

    index_of = {}
    for node in _elements(root, "bsdf"):
        ident = node.getAttribute("id")
        if not ident:
            raise SceneFormatError("BSDF must have an id")
        kind = node.getAttribute("type")
        if kind == "diffuse":
            rec = {"type": 0, "id": ident, "reflectance": _texture_value(_require(node, "reflectance", "reflectance"), 3, xml_dir)}
        elif kind == "roughconductor":
            alpha = _texture_value(_require(node, "alpha", "alpha"), 1, xml_dir)
            rec = {"type": 1, "id": ident, "alpha_u": alpha, "alpha_v": alpha.copy(),
                   "eta": _texture_value(_require(node, "eta", "eta"), 3, xml_dir), "k": _texture_value(_require(node, "k", "k"), 3, xml_dir)}
        else:
            raise SceneFormatError("Unsupported BSDF: " + kind)
        if ident in index_of:
            raise SceneFormatError("Duplicate BSDF id: " + ident)
        index_of[ident] = len(scene["bsdfs"])
        scene["bsdfs"].append(rec)

    for node in _elements(root, "emitter"):
        if node.getAttribute("type") != "envmap":
            raise SceneFormatError("Unsupported emitter: " + node.getAttribute("type"))
        if scene["envmap"] is not None:
            raise SceneFormatError("A scene is only allowed to have one envmap!")
        fn = _elements(node, "string")
        if not fn or fn[0].getAttribute("name") != "filename":
            raise SceneFormatError("Failed to retrieve bitmap filename")
        sc = _named(node, "scale")
        tr = _elements(node, "transform")
        scene["envmap"] = {"radiance": np.ascontiguousarray(read_exr(find_file(fn[0].getAttribute("value"), xml_dir))[:, :, :3]),
                           "scale": float(sc.getAttribute("value")) if sc is not None else 1.0, "to_world": Transform.from_xml(tr[0] if tr else None)}

    for node in _elements(root, "shape"):
        if node.getAttribute("type") != "obj":
            raise SceneFormatError("Unsupported shape: " + node.getAttribute("type"))
        fn = _elements(node, "string")
        if not fn or fn[0].getAttribute("name") != "filename":
            raise SceneFormatError("Missing mesh filename")
        mesh = read_obj(find_file(fn[0].getAttribute("value"), xml_dir))
        refs = _elements(node, "ref")
        if not refs:
            raise SceneFormatError("Missing BSDF reference")
        if _elements(node, "bsdf"):
            raise SceneFormatError("BSDFs declared under shapes are not supported.")
        if refs[0].getAttribute("id") not in index_of:
            raise SceneFormatError("Unknown BSDF id: " + refs[0].getAttribute("id"))
        flat = _named(node, "face_normals", "faceNormals")
        tr = _elements(node, "transform")
        mesh.update({"bsdf": index_of[refs[0].getAttribute("id")], "face_normals": flat is not None and flat.getAttribute("value") == "true",
                     "enable_edges": True, "id": node.getAttribute("id"), "to_world": Transform.from_xml(tr[0] if tr else None)})
        lights = _elements(node, "emitter")
        if lights:
            if lights[0].getAttribute("type") != "area":
                raise SceneFormatError("Unsupported emitter: " + lights[0].getAttribute("type"))
            rad = _texture_value(_require(lights[0], "radiance", "radiance"), 3, xml_dir).reshape(3)
            scene["emitters"].append({"mesh": len(scene["meshes"]), "radiance": rad})
        scene["meshes"].append(mesh)
    return scene
