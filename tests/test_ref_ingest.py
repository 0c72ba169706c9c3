"""Pins the ingest side of the path to the reference's OWN code: oracle/_ref/libref_ingest.so is the reference's vendored tinyobj /
tinyexr (+ miniz), compiled in place from /root/reference by oracle/build_ref.sh and called the way Mesh::load (mesh.cpp:62-141) and
BitmapLoader::load_openexr_rgba (bitmap_loader.cpp:13-53) call them. Every fixture OBJ and the environment map must come out of the
oracle's loader, the product's Python loader and the product's C++ loader exactly as out of the reference's parsers.
(The renderer itself is pinned separately, against its own source compiled with stand-ins for Enoki + OptiX: tests/test_ref_render.py.)"""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from conftest import ROOT

REF = os.path.join(ROOT, "oracle", "_ref", "libref_ingest.so")
OBJS = sorted(glob.glob(os.path.join(ROOT, "tests", "data", "objects", "*", "*.obj")))


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libref_ingest.so not built (needs /root/reference: bash oracle/build_ref.sh)")
    return C.CDLL(REF)


def ref_obj(ref, path):
    nv, nuv, nf = C.c_int(), C.c_int(), C.c_int()
    assert ref.ref_load_obj(path.encode(), C.byref(nv), C.byref(nuv), C.byref(nf)) == 0
    v = np.zeros((nv.value, 3), np.float32); uv = np.zeros((nuv.value, 2), np.float32)
    f = np.zeros((nf.value, 3), np.int32); uf = np.zeros((nf.value if nuv.value else 0, 3), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a.size else None
    ref.ref_get_obj(p(v), p(uv), p(f), p(uf))
    return v, uv, f, uf


@pytest.mark.parametrize("path", OBJS, ids=[os.path.relpath(p, os.path.join(ROOT, "tests", "data", "objects")) for p in OBJS])
def test_obj_loaders_match_the_references_tinyobj(ref, path):
    from oracle import orc
    from psdr_cuda_b200 import scene_io
    v, uv, f, uf = ref_obj(ref, path)
    assert len(v) > 0 and len(f) > 0
    for name, mesh in (("oracle", orc.load_obj(path)), ("product", scene_io.read_obj(path))):
        assert np.array_equal(mesh["verts"], v), name
        assert np.array_equal(mesh["faces"], f), name
        if len(uv):
            assert np.array_equal(mesh["uvs"], uv) and np.array_equal(mesh["uv_faces"], uf), name
        else:
            assert "uvs" not in mesh, name


def test_cpp_host_loader_matches_the_references_tinyobj(ref):
    import psdr_cuda_b200.compat  # noqa: F401
    import psdr_cuda
    for path in OBJS:
        v, uv, f, uf = ref_obj(ref, path)
        m = psdr_cuda.Mesh()
        m.load(path)
        assert np.array_equal(np.asarray(m.vertex_positions, np.float32), v), path
        assert np.array_equal(np.asarray(m.face_indices, np.int32), f), path
        if len(uv):
            assert np.array_equal(np.asarray(m.vertex_uv, np.float32), uv) and np.array_equal(np.asarray(m.face_uv_indices, np.int32), uf), path


def test_exr_decode_matches_the_references_tinyexr(ref):
    from oracle import orc
    from psdr_cuda_b200 import scene_io
    path = os.path.join(ROOT, "tests", "data", "envmaps", "ballroom_1k.exr")
    w, h = C.c_int(), C.c_int()
    assert ref.ref_load_exr(path.encode(), C.byref(w), C.byref(h)) == 0
    rgba = np.zeros((h.value, w.value, 4), np.float32)
    ref.ref_get_exr(rgba.ctypes.data_as(C.c_void_p))
    assert (w.value, h.value) == (1024, 512)
    for name, img in (("oracle", orc.load_exr(path)), ("product", scene_io.read_exr(path))):
        assert img.shape[:2] == (512, 1024), name
        assert np.array_equal(img[:, :, :3], rgba[:, :, :3]), name      # bit-identical half -> float decode
