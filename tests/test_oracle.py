"""CPU tests: the oracle against its golden vectors / known answers, and the physics of its derivatives."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, scene_path
from oracle import orc


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "cbox_bunny_golden.npz"))


def test_pcg32_canonical_kat():
    # pcg-random.org pcg32 demo, seed(42, 54): the generator enoki::PCG32 follows
    want = np.array([0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e], dtype=np.uint32)
    assert np.array_equal(orc.pcg32_kat(42, 54, 6), want)


def test_sampler_streams_kat():
    # src/core/sampler.cpp:8-40 (64-bit TEA quirk); values recorded in SURVEY §8c
    want = {0: ([0xca72a8ec, 0x0dfa513c, 0xd231a954], [0.79081202, 0.05460072, 0.82107019]),
            1: ([0x9747c25e, 0xf3522d60, 0x2540fe38], [0.59093869, 0.95047259, 0.14552295]),
            2: ([0x257788d7, 0x0b655a1c, 0x5f50fb14], [0.14635515, 0.04451525, 0.37232935])}
    for lane, (u, f) in want.items():
        gu, gf = orc.sampler_kat(lane, 3)
        assert np.array_equal(gu, np.array(u, dtype=np.uint32))
        assert np.allclose(gf, np.array(f, dtype=np.float32), atol=1e-8)


def test_golden_rng(golden):
    for lane in (0, 1, 2, 12345, 67108863):
        u, f = orc.sampler_kat(lane, 8)
        assert np.array_equal(u, golden["sampler_u32_%d" % lane])
        assert np.array_equal(f, golden["sampler_f32_%d" % lane])


def test_loader_fixture_sizes(cbox_desc):
    # SURVEY Appendix C
    assert len(cbox_desc["meshes"]) == 7 and len(cbox_desc["bsdfs"]) == 4 and len(cbox_desc["emitters"]) == 1
    assert sum(len(m["faces"]) for m in cbox_desc["meshes"]) == 69642
    assert cbox_desc["opts"] == dict(width=256, height=256, spp=8, sppe=8, sppse=8)
    # tinyobj fan triangulation of `f 4 3 2 1`
    assert cbox_desc["meshes"][0]["faces"].tolist() == [[3, 2, 1], [3, 1, 0]]
    assert np.allclose(cbox_desc["emitters"][0]["radiance"], [20, 20, 8])


def test_edge_lists(cbox_desc, golden):
    sc = orc.Scene(cbox_desc, dict(width=8, height=8, spp=1, sppe=0, sppse=0))
    sc.configure()
    counts = [len(sc.mesh_edges(m)) for m in range(7)]
    assert counts == golden["edges_counts"].tolist()
    assert sum(counts) == 104475
    e = sc.mesh_edges(1)
    assert np.array_equal(e[:64], golden["edges_bunny_head"])
    assert np.all(e[:, 0] < e[:, 1]) and np.all(e[:, 3] >= 0)   # bunny is closed 2-manifold
    order = np.lexsort((e[:, 1], e[:, 0]))
    assert np.array_equal(order, np.arange(len(e)))
    quad = sc.mesh_edges(0)
    assert len(quad) == 5 and (quad[:, 3] < 0).sum() == 4


def test_golden_tables_and_trace(cbox_desc, golden):
    sc = orc.Scene(cbox_desc, dict(width=32, height=32, spp=4, sppe=0, sppse=0))
    sc.configure()
    ti = sc.triangle_info()
    assert np.array_equal(ti[:8], golden["tri_info_first"]) and np.array_equal(ti[-8:], golden["tri_info_last"])
    assert np.allclose(ti.astype(np.float64).sum(axis=0), golden["tri_info_sum"], rtol=1e-12)
    tri, shape, u, v, t = sc.trace(golden["trace_o"], golden["trace_d"])
    assert np.array_equal(tri, golden["trace_tri"]) and np.array_equal(shape, golden["trace_shape"])
    assert np.array_equal(u, golden["trace_u"]) and np.array_equal(v, golden["trace_v"]) and np.array_equal(t, golden["trace_t"])
    sub = slice(0, 256)
    tb = sc.trace(golden["trace_o"][sub], golden["trace_d"][sub], brute=True)
    assert np.array_equal(tb[0], tri[sub]) and np.array_equal(tb[4], t[sub])
    # tmax culls: nothing is reported beyond tmax, hits closer than RayEpsilon are skipped
    tm = np.full(len(tri), 50.0, np.float32)
    tri2, _, _, _, t2 = sc.trace(golden["trace_o"], golden["trace_d"], tmax=tm)
    assert np.all((tri2 < 0) | (t2 < 50.0)) and np.all(t2[tri2 >= 0] > 1e-3)
    assert np.array_equal(tri2 >= 0, (tri >= 0) & (t < 50.0))


@pytest.mark.parametrize("name,make", [("direct11", lambda: orc.DirectIntegrator(1, 1)), ("direct21", lambda: orc.DirectIntegrator(2, 1)),
                                       ("path3", lambda: orc.PathIntegrator(3)), ("field_depth", lambda: orc.FieldExtractionIntegrator("depth")),
                                       ("field_shn", lambda: orc.FieldExtractionIntegrator("shNormal"))])
def test_golden_renderC(cbox_desc, golden, name, make):
    sc = orc.Scene(cbox_desc, dict(width=32, height=32, spp=4, sppe=0, sppse=0))
    sc.configure()
    integ = make()
    assert np.array_equal(integ.renderC(sc), golden["renderC_" + name])
    second = integ.renderC(sc)
    assert np.array_equal(second, golden["renderC2_" + name])
    if not name.startswith("field"):
        assert not np.array_equal(second, golden["renderC_" + name])     # streams persist across calls


def test_path_depth1_equals_direct11(cbox_desc):
    # the only reference pin a PathIntegrator can have
    opts = dict(width=24, height=24, spp=4, sppe=0, sppse=0)
    a = orc.Scene(cbox_desc, opts); a.configure()
    b = orc.Scene(cbox_desc, opts); b.configure()
    assert np.array_equal(orc.DirectIntegrator(1, 1).renderC(a), orc.PathIntegrator(1).renderC(b))
    a = orc.Scene(cbox_desc, opts); a.configure()
    b = orc.Scene(cbox_desc, opts); b.configure()
    ia, ib = orc.DirectIntegrator(1, 1).renderD(a), orc.PathIntegrator(1).renderD(b)
    assert np.array_equal(ia[0], ib[0])


def test_renderD_primal_matches_renderC_and_albedo_derivative_is_exact(cbox_desc, golden):
    opts = dict(width=32, height=32, spp=4, sppe=0, sppse=0)
    sc = orc.Scene(cbox_desc, opts)
    sc.set_bsdf_tangent(0, "reflectance", np.ones((1, 1, 3), np.float32))
    sc.configure()
    img, dimg = orc.PathIntegrator(3).renderD(sc)
    assert np.array_equal(img, golden["renderD_path3"]) and np.array_equal(dimg, golden["renderD_path3_dwhite"])
    ref = golden["renderC_path3"]
    err = np.abs(img - ref).mean(axis=1)
    assert np.mean(err > 1e-3) < 0.01                       # same estimator up to the AD formulation's rounding / edge flips
    # the image is a polynomial in the albedo: compare the tangent with a central difference of the same paths
    eps = 1e-2
    imgs = []
    for sgn in (+1, -1):
        s2 = orc.Scene(cbox_desc, opts)
        s2.set_bsdf_texture(0, "reflectance", np.full((1, 1, 3), 0.95 + sgn * eps, np.float32))
        s2.configure()
        imgs.append(orc.PathIntegrator(3).renderD(s2)[0])
    fd = (imgs[0] - imgs[1]) / (2 * eps)
    assert np.abs(fd - dimg).mean() < 2e-3 * max(1.0, np.abs(dimg).mean())


def test_discrete_distribution_semantics():
    # src/core/pmf.cpp:30-50: first i with cmf[i] >= u*sum; reuse rescales u into [0,1]; size-1 shortcut leaves u alone
    L = orc.lib()
    import ctypes as C
    pmf = np.array([1.0, 3.0, 0.0, 4.0], np.float32)
    u = np.array([0.0, 0.1249, 0.125, 0.3, 0.5, 0.50001, 0.9999], np.float32)
    idx = np.empty(len(u), np.int32); pdf = np.empty(len(u), np.float32); uo = np.empty(len(u), np.float32)
    L.orc_discrete_sample_reuse(pmf.ctypes.data_as(C.c_void_p), 4, u.ctypes.data_as(C.c_void_p), len(u), idx.ctypes.data_as(C.c_void_p),
                                pdf.ctypes.data_as(C.c_void_p), uo.ctypes.data_as(C.c_void_p))
    assert idx.tolist() == [0, 0, 0, 1, 1, 3, 3]
    assert np.allclose(pdf, [0.125, 0.125, 0.125, 0.375, 0.375, 0.5, 0.5])
    assert np.all((uo >= 0) & (uo <= 1)) and np.isclose(uo[3], (0.3 * 8 - 1) / 3, atol=1e-6)
    one = np.array([2.5], np.float32)
    L.orc_discrete_sample_reuse(one.ctypes.data_as(C.c_void_p), 1, u.ctypes.data_as(C.c_void_p), len(u), idx.ctypes.data_as(C.c_void_p),
                                pdf.ctypes.data_as(C.c_void_p), uo.ctypes.data_as(C.c_void_p))
    assert np.all(idx == 0) and np.all(pdf == 1) and np.array_equal(uo, u)


def test_unconfigured_scene_raises(cbox_desc):
    sc = orc.Scene(cbox_desc, dict(width=8, height=8, spp=1, sppe=0, sppse=0))
    with pytest.raises(RuntimeError, match="must be configured"):
        orc.DirectIntegrator(1, 1).renderC(sc)


def test_envmap_scale_tangent_is_the_exact_derivative():
    """The environment map's sampling distribution is normalised, so the image is linear in EnvironmentMap.scale:
    the forward-mode tangent must equal (I(1.1 s) - I(s)) / (0.1 s) at the same seeds."""
    opts = dict(width=24, height=24, spp=4, sppe=0, sppse=0)
    desc = orc.load_scene_description(scene_path("bunny_env"))
    s0 = orc.Scene(desc, opts)
    s0.set_envmap_tangent(None, 1.0)
    s0.configure()
    img, dimg = orc.DirectIntegrator(1, 1).renderD(s0)
    assert np.abs(dimg).max() > 0
    scale = desc["envmap"]["scale"]
    assert np.allclose(dimg * scale, img, rtol=2e-4, atol=1e-6)


def test_derivative_goldens():
    """forward-mode derivative images of every leaf kind (rough conductor, environment map, sensor pose, vertices) against the
    committed projections (tests/golden/derivative_golden.npz, generator make_golden_derivatives.py)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_derivatives", os.path.join(GOLDEN, "make_golden_derivatives.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(os.path.join(GOLDEN, "derivative_golden.npz"))
    got = mod.compute()
    assert set(got) == set(gold.files)
    for k, v in got.items():
        g = gold[k]
        assert np.all(np.isfinite(v)) and g[1] > 0, k
        # projected derivative (relative to the derivative image's L1 mass: OpenMP changes fp32 summation order), L1 mass, image sum
        assert abs(v[0] - g[0]) <= 1e-4 * g[1] and abs(v[1] - g[1]) <= 1e-4 * g[1] and abs(v[2] - g[2]) <= 1e-5 * abs(g[2]), (k, v, g)
