import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")

collect_ignore_glob = ["data/*"]   # fixtures (tests/data/examples/ holds the reference's own scripts, run by test_examples_verbatim.py)

DATA = os.path.join(ROOT, "tests", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def scene_path(name):
    return os.path.join(DATA, "scenes", name + ".xml")


@pytest.fixture(scope="session")
def cbox_desc():
    from oracle import orc
    return orc.load_scene_description(scene_path("cbox_bunny"))


@pytest.fixture(scope="session")
def native_lib():
    from psdr_cuda_b200 import build
    core = build.build_core()
    build.build_host()
    return core


# ---- a small scene with what the reference's example scenes lack: uv-mapped meshes, bitmap lookups, two sensors, mesh transforms with a rotation
TEXTURED_XML = """<?xml version='1.0' encoding='utf-8'?>
<scene version="0.5.0">
    <integrator type="direct"/>
    <sensor type="perspective">
        <float name="fov" value="40"/>
        <transform name="to_world"><lookat origin="0.3, 2.5, 4.0" target="0, 0.4, 0" up="0, 1, 0"/></transform>
        <sampler type="independent"><integer name="sampleCount" value="2"/></sampler>
        <film type="hdrfilm"><integer name="width" value="24"/><integer name="height" value="16"/></film>
    </sensor>
    <sensor type="perspective">
        <float name="fov" value="55"/>
        <transform name="to_world"><lookat origin="-2.0, 1.5, 3.0" target="0, 0.2, 0" up="0, 1, 0"/></transform>
    </sensor>
    <bsdf type="diffuse" id="tex"><rgb name="reflectance" value="0.5, 0.5, 0.5"/></bsdf>
    <bsdf type="roughconductor" id="metal">
        <float name="alpha" value="0.3"/><rgb name="eta" value="0.2, 0.9, 1.1"/><rgb name="k" value="3.9, 2.4, 2.2"/>
    </bsdf>
    <bsdf type="diffuse" id="black"><rgb name="reflectance" value="0, 0, 0"/></bsdf>
    <shape type="obj" id="floor"><string name="filename" value="%(dir)s/floor.obj"/><ref id="tex"/></shape>
    <shape type="obj" id="wall"><string name="filename" value="%(dir)s/wall.obj"/><ref id="metal"/>
        <transform name="to_world"><rotate y="1" angle="20"/><translate x="0.2" y="0" z="-1.2"/></transform></shape>
    <shape type="obj" id="light"><string name="filename" value="%(dir)s/light.obj"/><ref id="black"/>
        <emitter type="area"><rgb name="radiance" value="12, 10, 8"/></emitter></shape>
</scene>
"""


def _quad(path, corners, uvs=None):
    with open(path, "w") as fh:
        for c in corners:
            fh.write("v %g %g %g\n" % tuple(c))
        if uvs is not None:
            for t in uvs:
                fh.write("vt %g %g\n" % tuple(t))
            fh.write("f 1/1 2/2 3/3\nf 1/1 3/3 4/4\n")
        else:
            fh.write("f 1 2 3\nf 1 3 4\n")


@pytest.fixture(scope="session")
def textured_scene(tmp_path_factory):
    """two uv-mapped quads (a diffuse floor, a rough-conductor wall) under an area light, two sensors: the bitmap / uv / multi-sensor
    paths none of the reference's example scenes exercises"""
    d = tmp_path_factory.mktemp("textured")
    _quad(d / "floor.obj", [(-2, 0, 2), (2, 0, 2), (2, 0, -2), (-2, 0, -2)], [(0, 0), (1.7, 0), (1.7, 1.3), (0, 1.3)])
    _quad(d / "wall.obj", [(-1.5, 0, 0), (1.5, 0, 0), (1.5, 2, 0), (-1.5, 2, 0)], [(0.1, 0.2), (0.9, 0.1), (1.2, 0.8), (0.0, 1.0)])
    _quad(d / "light.obj", [(-0.6, 3, 0.6), (-0.6, 3, -0.6), (0.6, 3, -0.6), (0.6, 3, 0.6)])
    xml = d / "textured.xml"
    xml.write_text(TEXTURED_XML % dict(dir=str(d)))
    return str(xml)
