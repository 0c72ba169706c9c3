"""CPU tests of the drop-in boundary: the shared library builds, loads and exports every symbol include/psdr_b200.h
declares; the product refuses to run without a GPU (no CPU fallback); host-side loaders agree with the oracle's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, scene_path


def header_symbols():
    text = open(os.path.join(ROOT, "include", "psdr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native_lib):
    lib = C.CDLL(native_lib)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    assert lib.pb_version() >= 100


def test_python_binding_lists_the_same_symbols(native_lib):
    from psdr_cuda_b200 import capi
    assert sorted(capi.SYMBOLS) == header_symbols()


def test_header_documents_every_debug_key():
    # pb_debug_set's keys are the A/B switches behind profiles/: each one the library accepts is listed in the header
    src = open(os.path.join(ROOT, "psdr_cuda_b200", "csrc", "pb_capi.cu")).read()
    keys = set(re.findall(r'strcmp\(key, "([a-z0-9_]+)"\)', src))
    assert len(keys) >= 10
    header = open(os.path.join(ROOT, "include", "psdr_b200.h")).read()
    for k in keys:
        assert '"%s"' % k in header, "pb_debug_set key not documented in include/psdr_b200.h: " + k


def test_no_cpu_fallback(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from psdr_cuda_b200 import capi
    with pytest.raises(RuntimeError, match="CUDA device"):
        capi.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "psdr_cuda_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", "").replace("oracle's", "").replace("as the oracle", ""), f


@pytest.mark.parametrize("name", ["cbox_bunny", "cbox_bunny_mutiemitter", "bunny", "tree", "bunny_env", "bunny_env_2"])
def test_product_loader_matches_oracle_loader(name):
    from oracle import orc
    from psdr_cuda_b200 import scene_io
    a, b = orc.load_scene_description(scene_path(name)), scene_io.load_scene_description(scene_path(name))
    assert a["opts"] == b["opts"] and len(a["meshes"]) == len(b["meshes"]) and len(a["bsdfs"]) == len(b["bsdfs"])
    for ma, mb in zip(a["meshes"], b["meshes"]):
        for k in ("verts", "faces", "to_world"):
            assert np.array_equal(ma[k], mb[k])
        assert ma["bsdf"] == mb["bsdf"] and ma["face_normals"] == mb["face_normals"] and ("uvs" in ma) == ("uvs" in mb)
    for sa, sb in zip(a["sensors"], b["sensors"]):
        assert np.array_equal(sa["to_world"], sb["to_world"]) and sa["fov"] == sb["fov"] and sa["near"] == sb["near"] and sa["far"] == sb["far"]
    assert [e["mesh"] for e in a["emitters"]] == [e["mesh"] for e in b["emitters"]]
    if a["envmap"] is not None:
        assert np.array_equal(a["envmap"]["radiance"], b["envmap"]["radiance"]) and a["envmap"]["scale"] == b["envmap"]["scale"]


def test_loader_errors():
    from psdr_cuda_b200 import scene_io
    with pytest.raises(RuntimeError, match="XML parsing failed"):
        scene_io.load_scene_description(xml_string="<scene")
    bad = "<scene><sensor type='perspective'><float name='fov' value='30'/><film><integer name='width' value='4'/><integer name='height' value='4'/></film></sensor></scene>"
    with pytest.raises(RuntimeError, match="Missing sampler node"):
        scene_io.load_scene_description(xml_string=bad)
    bad2 = "<scene><bsdf type='plastic' id='x'/></scene>"
    with pytest.raises(RuntimeError, match="Unsupported BSDF"):
        scene_io.load_scene_description(xml_string=bad2)
