"""Pins the oracle to psdr-cuda's OWN renderer source, run here on the CPU: oracle/_ref/libref_render.so is every src/**/*.cpp of the
reference on the rendering path (scene loader, Scene::configure, Mesh, PerspectiveCamera, BSDFs, emitters, distributions, sampler,
Integrator::renderC / renderD with the primary- and secondary-edge terms, DirectIntegrator incl. edge guiding, FieldExtractionIntegrator)
compiled UNMODIFIED from /root/reference (oracle/build_ref.sh) against two stand-ins for its external dependencies: oracle/ref_dyn (Enoki:
host arrays, forward-mode tangents in place of the tape) and the Scene_OptiX in oracle/ref_render_shim.cpp (exact closest hit with the
reference's own ray/triangle arithmetic). Every test runs the same steps through that library and through the oracle.

What agrees and how closely: integer tables exactly; fp32 tables to the last bit or two; images and forward-mode derivative images to
~1e-5 of the image maximum per pixel, EXCEPT a small fraction of pixels (bounded below) in which one lane took the other side of a
knife-edge decision — a shadow ray leaving a surface at grazing incidence, a primary-edge ray 1e-5 (sample space) off a silhouette — because
the two differ in last bits (the path-space hit point of the D flavour, the 4x4 inverse behind world_to_sample). Those lanes are
all-or-nothing in both directions (seen lane by lane in test_lane_radiance); they are fp32 conditioning of the reference's algorithm, not
a difference in it.

What stays assumed: Enoki's and OptiX's own semantics — the stand-ins implement the same assumptions the oracle makes
(exact rcp / rsqrt, sequential sums, closest hit with ties to the lowest id), so this file pins the oracle's reading of psdr-cuda's code,
statement by statement, not Enoki's rounding."""
import os

import numpy as np
import pytest

from conftest import ROOT, scene_path
from oracle import orc, refrun

pytestmark = pytest.mark.skipif(not refrun.available(), reason="oracle/_ref/libref_render.so not built and /root/reference absent")
TESTS = os.path.join(ROOT, "tests")
SCENES = ["bunny", "bunny_env", "bunny_env_2", "cbox_bunny", "cbox_bunny_mutiemitter", "cbox_bunny_rc", "tree"]
_desc = {}


def desc(name):
    if name not in _desc:
        _desc[name] = orc.load_scene_description(scene_path(name))
    return _desc[name]


def pair(name, w, h, spp, sppe=0, sppse=0, plain=True, configure=True):
    """the same scene through the reference's loader + Scene and through the oracle's"""
    refrun.set_matvec_plain(plain)
    r = refrun.Scene(scene_path(name), TESTS, w, h, spp, sppe, sppse)
    o = orc.Scene(desc(name), dict(width=w, height=h, spp=spp, sppe=sppe, sppse=sppse))
    if configure:
        r.configure()
        o.configure()
    return r, o


def assert_images_close(a, b, rel=2e-4, outliers=0.0, what=""):
    """a ~ b per pixel relative to the image maximum, except for at most `outliers` (fraction of the non-zero pixels, at least one pixel
    allowed when > 0): the knife-edge lanes described in the module docstring"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.isfinite(a).all() and np.isfinite(b).all(), what
    scale = max(np.abs(b).max(), 1e-12)
    d = np.abs(a - b).max(axis=1)
    bad = int((d > rel * scale).sum())
    nz = int(((np.abs(a).max(axis=1) > 0) | (np.abs(b).max(axis=1) > 0)).sum())
    allowed = 0 if outliers == 0 else max(1, int(np.ceil(outliers * nz)))
    assert bad <= allowed, "%s: %d of %d non-zero pixels differ by more than %g of the maximum (allowed %d); worst %g" % (what, bad, nz, rel, allowed, d.max() / scale)
    return bad


# ---- Scene::configure: tables -------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SCENES)
def test_tables_match_reference_source(name):
    """mesh.cpp:19-51,215-274 (triangle info, vertex normals, secondary edges), mesh.cpp:143-203 (edge list), perspective.cpp:11-111
    (camera matrices, primary-edge list), scene.cpp:56-278 (global tables, bounding mesh for the envmap)"""
    r, o = pair(name, 20, 20, 2, 2, 2)
    nm = r.num_meshes()
    for m in range(nm - (1 if desc(name)["envmap"] is not None else 0)):   # the bounding mesh has no edge list
        assert np.array_equal(r.mesh_edges(m), o.mesh_edges(m)), (name, m)
    tr, to = r.triangle_info(), o.triangle_info()
    assert tr.shape == to.shape
    assert np.abs(tr - to).max() <= 4e-7 * max(1.0, np.abs(to).max()), np.abs(tr - to).max()
    sr, so = r.sensor_info(), o.sensor_info()
    for k in sr:
        assert np.allclose(sr[k], so[k], rtol=2e-6, atol=1e-6 * np.abs(so[k]).max()), (k, sr[k], so[k])
    er, eo = r.sec_edges(), o.sec_edges()
    pr, po = r.primary_edges(), o.primary_edges()
    assert er.shape == eo.shape and np.abs(er - eo).max() <= 4e-7 * max(1.0, np.abs(eo).max())
    assert pr.shape == po.shape
    # end points in sample space; the unit normal of an edge of length L carries the end points' rounding amplified by 1 / L
    assert (np.abs(pr[:, :4] - po[:, :4]) <= 2e-6 * np.maximum(1.0, np.abs(po[:, :4]))).all()
    mag = np.maximum(1.0, np.abs(po[:, :4]).max(axis=1))
    assert (np.abs(pr[:, 4:6] - po[:, 4:6]).max(axis=1) <= 1e-5 + 4e-7 * mag / np.maximum(po[:, 6], 1e-12)).all()
    assert np.allclose(pr[:, 6], po[:, 6], rtol=1e-3, atol=2e-6)


def test_with_the_forms_assumed_for_enoki():
    """the stand-in's two arithmetic switches set to what Enoki itself is believed to do — matrix * vector as an fmadd chain over columns, dot
    accumulated from the first component up — where the oracle and the CUDA product use plain sums and accumulate from the last component
    down: last bits of the tables (a few coplanar edges then fall on the other side of the 1 - EdgeEpsilon test, mesh.cpp:262), and an image
    within the usual tolerance"""
    refrun.set_dot_from_first(True)
    try:
        r, o = pair("cbox_bunny", 24, 24, 2, 1, 1, plain=False)
        tr, to = r.triangle_info(), o.triangle_info()
        assert np.abs(tr - to).max() <= 1e-6 * np.abs(to).max()
        assert abs(len(r.sec_edges()) - len(o.sec_edges())) <= 8
        a, b = refrun.DirectIntegrator(1, 1).renderC(r), orc.DirectIntegrator(1, 1).renderC(o)
        assert_images_close(a, b, rel=2e-4, outliers=0.01, what="renderC")
    finally:
        refrun.set_dot_from_first(False)
        refrun.set_matvec_plain(True)


def test_closest_hits_match():
    """Scene::ray_intersect<false> through the stand-in for OptiX against the oracle's tracer: same triangle, same barycentrics"""
    r, o = pair("cbox_bunny", 8, 8, 1)
    ti = o.triangle_info()
    rng = np.random.default_rng(1)
    n = 4000
    k = rng.integers(0, len(ti), n)
    uv = rng.uniform(0, 1, (n, 2))
    f = uv.sum(1) > 1
    uv[f] = 1 - uv[f]
    org = (ti[k, 0:3] + ti[k, 3:6] * uv[:, :1] + ti[k, 6:9] * uv[:, 1:]).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    tri_r, shape_r, u_r, v_r, t_r = r.trace(org, d)
    tri_o, shape_o, u_o, v_o, t_o = o.trace(org, d)
    assert np.array_equal(tri_r, tri_o) and np.array_equal(shape_r, shape_o)
    assert np.array_equal(u_r, u_o) and np.array_equal(v_r, v_o)
    hit = tri_o >= 0
    assert hit.sum() > 3000 and np.allclose(t_r[hit], t_o[hit], rtol=1e-4, atol=2e-3)   # its.t = |p - o| (scene.cpp:332) vs the hit distance


# ---- renderC ---------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,bsdf_samples,light_samples", [
    ("cbox_bunny", 1, 1), ("cbox_bunny", 2, 0), ("cbox_bunny", 0, 2), ("cbox_bunny", 2, 1),
    ("cbox_bunny_mutiemitter", 1, 1), ("cbox_bunny_rc", 1, 1), ("bunny_env", 1, 1), ("bunny_env_2", 1, 1), ("bunny_env_2", 0, 2), ("tree", 1, 1)])
def test_renderC_matches_reference_source(name, bsdf_samples, light_samples):
    """integrator.cpp:13-29,64-95 + direct.cpp:47-163 and everything below (samplers, camera rays, BSDF / emitter sampling, MIS), twice in a
    row: the second render continues the sampler streams (scene.cpp:65-79 reseeds only when the sample count changes)"""
    r, o = pair(name, 24, 24, 2)
    ri, oi = refrun.DirectIntegrator(bsdf_samples, light_samples), orc.DirectIntegrator(bsdf_samples, light_samples)
    for rep in range(2):
        a, b = ri.renderC(r), oi.renderC(o)
        assert np.abs(b).max() > 0
        assert_images_close(a, b, rel=2e-4, outliers=0.01, what="%s pass %d" % (name, rep))


def test_hide_emitters_and_field_integrators():
    """direct.cpp:51 (m_hide_emitters) and field.cpp:34-54 for every field"""
    r, o = pair("cbox_bunny", 24, 24, 2)
    assert_images_close(refrun.DirectIntegrator(1, 1, True).renderC(r), orc.DirectIntegrator(1, 1, True).renderC(o), outliers=0.01, what="hide_emitters")
    for field in ("silhouette", "position", "depth", "geoNormal", "shNormal", "uv"):
        a, b = refrun.FieldExtractionIntegrator(field).renderC(r), orc.FieldExtractionIntegrator(field).renderC(o)
        assert_images_close(a, b, rel=2e-5, what=field)


def test_missing_emitter_raises_like_the_reference():
    """scene.cpp:404: emitter sampling on a scene without emitters is an error ("No Emitter!"), BSDF sampling alone renders black"""
    r, o = pair("bunny", 12, 12, 1)
    with pytest.raises(RuntimeError, match="No Emitter"):
        refrun.DirectIntegrator(1, 1).renderC(r)
    with pytest.raises(RuntimeError, match="No Emitter"):
        orc.DirectIntegrator(1, 1).renderC(o)
    a, b = refrun.DirectIntegrator(1, 0).renderC(r), orc.DirectIntegrator(1, 0).renderC(o)
    assert np.abs(a).max() == 0 and np.abs(b).max() == 0


def test_lane_radiance():
    """lane by lane, before the scatter to pixels: the lanes agree to fp32 rounding except a few per ten thousand that are all-or-nothing
    (one side a full contribution, the other exactly zero) — the knife-edge lanes the image tolerances above allow for"""
    r, o = pair("cbox_bunny", 32, 32, 4)
    ri, oi = refrun.DirectIntegrator(1, 1), orc.DirectIntegrator(1, 1)
    n = 32 * 32 * 4
    L = orc.lib()
    for ad in (0, 1):
        if ad:
            r, o = pair("cbox_bunny", 32, 32, 4)
        a = ri.lane_radiance(r, ad=bool(ad))
        b = np.zeros((n, 3), np.float32)
        o3 = np.zeros(3, np.float32)
        for lane in range(n):
            L.orc_debug_lane(o.h, oi.h, 0, orc.C.c_int64(lane), ad, o3.ctypes.data_as(orc.C.c_void_p))
            b[lane] = o3
        d = np.abs(a - b).max(axis=1)
        scale = np.maximum(np.abs(b).max(axis=1), 1e-2)
        bad = np.nonzero(d > 1e-3 * scale)[0]
        assert len(bad) <= 4, (ad, len(bad))
        for lane in bad:
            assert np.abs(a[lane]).max() == 0 or np.abs(b[lane]).max() == 0, (lane, a[lane], b[lane])


# ---- renderD: forward-mode derivative images ---------------------------------------------------------------------------------------------------
def _seed(r, o, leaf, rng, name):
    """set the same tangent on the same leaf in both; returns whether configure() has to run afterwards"""
    kind = leaf[0]
    if kind == "bsdf":
        _, b, tex, t = leaf
        t = np.asarray(t, np.float32).reshape(1, 1, -1)
        r.set_bsdf_tangent(b, tex, t)
        o.set_bsdf_tangent(b, tex, t)
    elif kind == "vertices":
        t = rng.normal(size=(r.num_vertices(leaf[1]), 3)).astype(np.float32)
        r.set_mesh_vertex_tangent(leaf[1], t)
        o.set_mesh_vertex_tangent(leaf[1], t)
    elif kind == "translate":
        t = np.zeros((r.num_vertices(leaf[1]), 3), np.float32)
        t[:] = np.asarray(leaf[2], np.float32)
        r.set_mesh_vertex_tangent(leaf[1], t)
        o.set_mesh_vertex_tangent(leaf[1], t)
    elif kind == "mesh_transform":
        t = np.zeros((4, 4), np.float32)
        t[:3, 3] = [1.0, 0.5, -0.3]
        t[0, 1], t[2, 0] = 0.1, -0.05
        r.set_mesh_transform_tangent(leaf[1], t, leaf[2])
        o.set_mesh_transform_tangent(leaf[1], t, leaf[2])
    elif kind == "sensor":
        A = rng.normal(size=(3, 3))
        t = np.zeros((4, 4), np.float32)
        t[:3, :3] = (0.01 * (A - A.T)) @ desc(name)["sensors"][0]["to_world"][:3, :3]
        t[:3, 3] = [0.3, -0.2, 0.1]
        r.set_sensor_transform_tangent(0, t)
        o.set_sensor_transform_tangent(0, t)
    elif kind == "env_scale":
        r.set_envmap_tangent(None, 1.0)
        o.set_envmap_tangent(None, 1.0)
    elif kind == "env_radiance":
        t = rng.uniform(0, 1, desc(name)["envmap"]["radiance"].shape).astype(np.float32)
        r.set_envmap_tangent(t, 0.0)
        o.set_envmap_tangent(t, 0.0)
    elif kind == "env_transform":
        A = rng.normal(size=(3, 3))
        t = np.zeros((4, 4), np.float32)
        t[:3, :3] = 0.1 * (A - A.T)
        r.set_envmap_transform_tangent(t)
        o.set_envmap_transform_tangent(t)
    else:
        raise AssertionError(kind)


INTERIOR = [
    ("cbox_bunny", ("bsdf", 0, "reflectance", [1.0, 0.5, 0.25])),
    ("cbox_bunny", ("bsdf", 1, "reflectance", [0.3, 1.0, 0.6])),
    ("cbox_bunny", ("vertices", 1)),
    ("cbox_bunny", ("vertices", 0)),            # the emitter's vertices: light sampling, pdfs, J
    ("cbox_bunny", ("mesh_transform", 1, True)),
    ("cbox_bunny", ("mesh_transform", 1, False)),
    ("cbox_bunny", ("sensor",)),
    ("cbox_bunny_mutiemitter", ("vertices", 1)),
    ("cbox_bunny_rc", ("vertices", 1)),
    ("cbox_bunny_rc", ("bsdf", 3, "alpha_u", [1.0])),
    ("bunny_env", ("bsdf", 0, "alpha_u", [1.0])),
    ("bunny_env", ("bsdf", 0, "alpha_v", [1.0])),
    ("bunny_env", ("bsdf", 0, "eta", [1.0, 0.5, 0.25])),
    ("bunny_env", ("bsdf", 0, "k", [1.0, 0.5, 0.25])),
    ("bunny_env", ("bsdf", 0, "specular_reflectance", [1.0, 0.5, 0.25])),
    ("bunny_env", ("vertices", 0)),
    ("bunny_env", ("env_scale",)),
    ("bunny_env", ("env_radiance",)),
    ("bunny_env", ("env_transform",)),
    ("bunny_env_2", ("sensor",)),
]


@pytest.mark.parametrize("name,leaf", INTERIOR, ids=["%s-%s" % (n, "-".join(str(x) for x in l[:3] if not isinstance(x, list))) for n, l in INTERIOR])
def test_renderD_interior_derivatives_match_reference_source(name, leaf):
    """the `D` flavour of the whole interior path (integrator.cpp:64-95, direct.cpp:47-163 with its detach() calls, scene.cpp:281-381
    incl. the material-form Jacobians) differentiated in forward mode through the reference's own statements, per leaf kind"""
    rng = np.random.default_rng(7)
    r, o = pair(name, 24, 24, 2, configure=False)
    _seed(r, o, leaf, rng, name)
    r.configure()
    o.configure()
    (a, at), (b, bt) = refrun.DirectIntegrator(1, 1).renderD(r), orc.DirectIntegrator(1, 1).renderD(o)
    assert np.abs(bt).max() > 0
    assert_images_close(a, b, rel=2e-4, outliers=0.01, what="primal")
    assert_images_close(at, bt, rel=5e-4, outliers=0.02, what="tangent")


BOUNDARY = [
    # (scene, res, sppe, sppse, leaf, allowed outlier fraction): cbox_bunny's camera stands 1000 units away behind a 13-degree lens, so a
    # primary-edge ray pair 1e-5 apart in sample space is ~2e-3 world units from the silhouette, a few dozen ulps of the hit arithmetic
    ("bunny_env", 32, 8, 0, ("vertices", 0), 0.02),
    ("bunny_env", 32, 8, 0, ("sensor",), 0.02),
    ("cbox_bunny", 32, 16, 0, ("translate", 1, [1.0, 0.5, -0.3]), 0.2),
    ("cbox_bunny", 24, 0, 64, ("translate", 1, [1.0, 0.5, -0.3]), 0.05),
    ("cbox_bunny", 24, 0, 64, ("vertices", 1), 0.05),
    ("cbox_bunny_mutiemitter", 24, 0, 64, ("vertices", 1), 0.05),
    # one BSDF + environment map: the reference shades boundary segments that end on the envmap's bounding mesh with meshes[0]'s BSDF
    # (direct.cpp:278-284) — found by this test, reproduced in the oracle and the CUDA kernel
    ("bunny_env", 24, 0, 32, ("vertices", 0), 0.05),
    ("bunny_env_2", 24, 0, 32, ("sensor",), 0.05),
]


@pytest.mark.parametrize("name,res,sppe,sppse,leaf,outliers", BOUNDARY, ids=["%s-e%d-s%d-%s" % (b[0], b[2], b[3], b[4][0]) for b in BOUNDARY])
def test_renderD_boundary_terms_match_reference_source(name, res, sppe, sppse, leaf, outliers):
    """primary edges (integrator.cpp:98-119, perspective.cpp:139-200) and secondary edges (direct.cpp:207-316, scene.cpp:456-492) alone
    (spp = 0): edge sampling, the two-sided radiance difference, the normal velocity through ray_intersect_triangle<true>"""
    rng = np.random.default_rng(11)
    r, o = pair(name, res, res, 0, sppe, sppse, configure=False)
    _seed(r, o, leaf, rng, name)
    r.configure()
    o.configure()
    (a, at), (b, bt) = refrun.DirectIntegrator(1, 1).renderD(r), orc.DirectIntegrator(1, 1).renderD(o)
    assert np.abs(a).max() == 0 and np.abs(b).max() == 0      # value - detach(value): the boundary terms carry derivatives only
    assert np.abs(bt).max() > 0
    bad = assert_images_close(at, bt, rel=1e-3, outliers=outliers, what="tangent")
    # and as a whole: the images' projections agree to the share the outlier pixels can carry
    assert abs(at.sum() - bt.sum()) <= 1e-3 * np.abs(bt).sum() + (bad + 1) * np.abs(bt).max()


def test_all_terms_together_and_edge_guiding():
    """renderD with interior + both boundary terms after DirectIntegrator::preprocess_secondary_edges (direct.cpp:166-204: the guiding
    grid built from eval_secondary_edge<false>, then sample_reuse in render_secondary_edges, direct.cpp:207-222)"""
    rng = np.random.default_rng(3)
    r, o = pair("cbox_bunny", 24, 24, 2, 2, 32, configure=False)
    _seed(r, o, ("vertices", 1), rng, "cbox_bunny")
    r.configure()
    o.configure()
    ri, oi = refrun.DirectIntegrator(1, 1), orc.DirectIntegrator(1, 1)
    reso = [8, 4, 4, 8]
    ri.preprocess_secondary_edges(r, 0, reso, 2)
    oi.preprocess_secondary_edges(o, 0, reso, 2)
    (a, at), (b, bt) = ri.renderD(r), oi.renderD(o)
    assert_images_close(a, b, rel=2e-4, outliers=0.01, what="primal")
    assert_images_close(at, bt, rel=1e-3, outliers=0.1, what="tangent")


def test_bitmap_textures():
    """bitmap.cpp:56-96 (bilinear lookup with the flipped v, clamping at the last texel) under a textured diffuse BSDF: the texel derivative
    image; cbox walls carry no uv, so the lookup runs at uv = 0 there, the bunny has none either: use the textured quads of bunny_env_2 if
    it has uvs, else this degenerates to the constant-uv corner of the bilinear lookup, still through the reference's code"""
    rng = np.random.default_rng(5)
    r, o = pair("cbox_bunny", 20, 20, 2, configure=False)
    tex = rng.uniform(0.1, 0.9, (5, 7, 3)).astype(np.float32)
    tang = rng.uniform(0, 1, (5, 7, 3)).astype(np.float32)
    r.set_bsdf_texture(0, "reflectance", tex)
    o.set_bsdf_texture(0, "reflectance", tex)
    r.set_bsdf_tangent(0, "reflectance", tang.reshape(-1, 3))
    o.set_bsdf_tangent(0, "reflectance", tang)
    r.configure()
    o.configure()
    (a, at), (b, bt) = refrun.DirectIntegrator(1, 1).renderD(r), orc.DirectIntegrator(1, 1).renderD(o)
    assert_images_close(a, b, rel=2e-4, outliers=0.01, what="primal")
    assert_images_close(at, bt, rel=5e-4, outliers=0.02, what="tangent")


def test_stand_in_tangents_are_derivatives_of_the_reference_run():
    """independent of the oracle: the forward-mode tangent the Enoki stand-in carries through the reference's renderD equals the central
    finite difference of the reference's own renderD image, for parameters that do not steer the sampling (conductor eta / k enter only
    the Fresnel term, roughconductor.cpp:52-54; the envmap scale only the emitted radiance, envmap.cpp:53-57), with the sampler streams
    held fixed (a fresh scene per evaluation)"""
    def render(eta=None, k=None, tangent=None):
        refrun.set_matvec_plain(True)
        sc = refrun.Scene(scene_path("bunny_env"), TESTS, 16, 16, 2, 0, 0)
        if eta is not None:
            sc.set_bsdf_texture(0, "eta", np.asarray(eta, np.float32).reshape(1, 1, 3))
        if k is not None:
            sc.set_bsdf_texture(0, "k", np.asarray(k, np.float32).reshape(1, 1, 3))
        if tangent is not None:
            sc.set_bsdf_tangent(0, tangent[0], np.asarray(tangent[1], np.float32).reshape(1, 3))
        sc.configure()
        return refrun.DirectIntegrator(1, 1).renderD(sc)
    b = desc("bunny_env")["bsdfs"][0]
    eta0, k0 = b["eta"].reshape(3).astype(np.float64), b["k"].reshape(3).astype(np.float64)
    for name, base, direction in (("eta", eta0, np.array([1.0, 0.5, 0.25])), ("k", k0, np.array([0.3, 1.0, 0.6]))):
        _, dimg = render(tangent=(name, direction))
        h = 2e-2
        plus = render(**{name: base + h * direction})[0].astype(np.float64)
        minus = render(**{name: base - h * direction})[0].astype(np.float64)
        fd = (plus - minus) / (2 * h)
        assert np.abs(dimg).max() > 0
        assert np.abs(fd - dimg).max() <= 2e-3 * np.abs(dimg).max(), (name, np.abs(fd - dimg).max(), np.abs(dimg).max())


@pytest.mark.parametrize("sensor,leaf", [(0, "texels"), (1, "texels"), (0, "rc_texels"), (0, "vertices"), (1, "uv"), (0, "boundary")])
def test_textured_uv_mapped_scene_matches_reference_source(textured_scene, sensor, leaf):
    """bitmap.cpp:56-96 inside a render (bilinear texels at interpolated, wrapped uvs), their derivative w.r.t. texels, w.r.t. the vertices
    through the barycentrics (scene.cpp:338-341,361-364) and w.r.t. Mesh.vertex_uv, for a diffuse and a rough-conductor texture, from either
    of two sensors (scene_loader.cpp:246-292: film and sampler come from the first)"""
    rng = np.random.default_rng(21)
    refrun.set_matvec_plain(True)
    sppe = sppse = 8 if leaf == "boundary" else 0
    spp = 0 if leaf == "boundary" else 4
    r = refrun.Scene(textured_scene, TESTS, 0, 0, spp, sppe, sppse)
    d = orc.load_scene_description(textured_scene)
    assert (r.opts["width"], r.opts["height"]) == (24, 16) == (d["opts"]["width"], d["opts"]["height"])
    o = orc.Scene(d, dict(spp=spp, sppe=sppe, sppse=sppse))
    tex3 = rng.uniform(0.05, 0.95, (6, 9, 3)).astype(np.float32)
    tex1 = rng.uniform(0.1, 0.6, (5, 4, 1)).astype(np.float32)
    for s in (r, o):
        s.set_bsdf_texture(0, "reflectance", tex3)
        s.set_bsdf_texture(1, "alpha_u", tex1)
        s.set_bsdf_texture(1, "alpha_v", tex1)
    if leaf == "texels":
        t = rng.uniform(0, 1, tex3.shape).astype(np.float32)
        r.set_bsdf_tangent(0, "reflectance", t.reshape(-1, 3))
        o.set_bsdf_tangent(0, "reflectance", t)
    elif leaf == "rc_texels":
        t = rng.uniform(0, 1, tex1.shape).astype(np.float32)
        r.set_bsdf_tangent(1, "alpha_u", t.reshape(-1))
        o.set_bsdf_tangent(1, "alpha_u", t)
    elif leaf in ("vertices", "boundary"):
        for m in (0, 1):
            t = rng.normal(size=(4, 3)).astype(np.float32)
            r.set_mesh_vertex_tangent(m, t)
            o.set_mesh_vertex_tangent(m, t)
    elif leaf == "uv":
        t = rng.normal(size=(4, 2)).astype(np.float32)
        r.set_mesh_uv_tangent(0, t)
        o.set_mesh_uv_tangent(0, t)
    r.configure()
    o.configure()
    (a, at), (b, bt) = refrun.DirectIntegrator(1, 1).renderD(r, sensor), orc.DirectIntegrator(1, 1).renderD(o, sensor)
    assert np.abs(bt).max() > 0
    if leaf != "boundary":
        assert_images_close(a, b, rel=2e-4, outliers=0.01, what="primal")
        a2, b2 = refrun.DirectIntegrator(1, 1).renderC(r, sensor), orc.DirectIntegrator(1, 1).renderC(o, sensor)
        assert np.abs(b2).max() > 0
        assert_images_close(a2, b2, rel=2e-4, outliers=0.01, what="renderC")
    assert_images_close(at, bt, rel=1e-3, outliers=0.05 if leaf == "boundary" else 0.02, what="tangent")


@pytest.mark.parametrize("name,w,h,spp,sppe,sppse,bs,ls,hide,mesh", [
    ("cbox_bunny_mutiemitter", 30, 20, 1, 0, 0, 2, 3, False, 1),      # spp = 1 (no division of the lane index), several samples of each kind
    ("cbox_bunny_mutiemitter", 30, 20, 3, 2, 8, 3, 2, True, 0),       # hidden emitters, emitter-vertex tangent, all three terms
    ("tree", 24, 24, 2, 4, 16, 1, 1, False, 2),                        # rotated / scaled meshes, all three terms
    ("bunny_env_2", 32, 18, 2, 4, 16, 1, 1, False, 1),                 # two BSDFs + envmap: bounding-mesh end points get no BSDF
    ("cbox_bunny_rc", 20, 20, 2, 2, 8, 2, 2, False, 2),
])
def test_mixed_configurations_match_reference_source(name, w, h, spp, sppe, sppse, bs, ls, hide, mesh):
    """non-square films, spp = 1, several BSDF / emitter samples, hidden emitters, all three terms in one renderD"""
    rng = np.random.default_rng(5)
    r, o = pair(name, w, h, spp, sppe, sppse, configure=False)
    _seed(r, o, ("vertices", mesh), rng, name)
    r.configure()
    o.configure()
    (a, at), (b, bt) = refrun.DirectIntegrator(bs, ls, hide).renderD(r), orc.DirectIntegrator(bs, ls, hide).renderD(o)
    assert np.abs(b).max() > 0 and np.abs(bt).max() > 0
    assert_images_close(a, b, rel=1e-3, outliers=0.01, what="primal")
    assert_images_close(at, bt, rel=1e-3, outliers=0.03, what="tangent")


def test_zero_tangent_through_an_infinite_local_derivative_stays_zero():
    """camera rays that leave through a pole of the environment map clamp the latitude cosine to +-1, where acos has an infinite derivative;
    with a zero direction tangent the derivative image must stay finite (Enoki's autodiff multiplies with 0 * inf = 0; found with the oracle
    returning NaN in 2 pixels of this image)"""
    r, o = pair("bunny_env", 128, 128, 8, configure=False)
    t = np.array([[0.3, -0.2, 0.5]], np.float32)
    r.set_bsdf_tangent(0, "eta", t)
    o.set_bsdf_tangent(0, "eta", t.reshape(1, 1, 3))
    r.configure()
    o.configure()
    (a, at), (b, bt) = refrun.DirectIntegrator(1, 1).renderD(r), orc.DirectIntegrator(1, 1).renderD(o)
    assert np.isfinite(at).all() and np.isfinite(bt).all()
    for px in (9786, 10295):   # the two pole pixels
        assert np.abs(at[px] - bt[px]).max() <= 1e-3 * np.abs(bt).max()
    assert_images_close(a, b, rel=2e-4, outliers=0.01, what="primal")
    assert_images_close(at, bt, rel=1e-3, outliers=0.02, what="tangent")


@pytest.mark.parametrize("field", ["silhouette", "position", "depth", "geoNormal", "shNormal"])
def test_field_integrator_derivatives_match_reference_source(field):
    """FieldExtractionIntegrator in its D flavour (field.cpp:34-54 on the solid-angle intersection, scene.cpp:345-372) plus its primary-edge
    term (integrator.cpp:98-119) — the reference's examples/config.py `bunny_silhouette` setup: vertex tangent of the bunny"""
    rng = np.random.default_rng(3)
    r, o = pair("bunny", 32, 32, 2, 8, 0, configure=False)
    _seed(r, o, ("vertices", 0), rng, "bunny")
    r.configure()
    o.configure()
    (a, at), (b, bt) = refrun.FieldExtractionIntegrator(field).renderD(r), orc.FieldExtractionIntegrator(field).renderD(o)
    assert np.abs(bt).max() > 0
    assert_images_close(a, b, rel=2e-5, what="primal")
    assert_images_close(at, bt, rel=1e-3, outliers=0.02, what="tangent")


def test_xml_bitmap_textures_and_transformed_envmap_render(tmp_path, monkeypatch):
    """a scene whose XML loads OpenEXR bitmaps (a diffuse reflectance and a rough-conductor roughness texture; the meshes carry no uvs, so
    the lookups run at uv = 0) under an environment map with a scale and a rotated to_world: loader -> configure -> renderC / renderD"""
    from test_host_module import LOADER_CASES
    xml = tmp_path / "textures.xml"
    xml.write_text(LOADER_CASES["textures"])
    monkeypatch.chdir(TESTS)
    refrun.set_matvec_plain(True)
    r = refrun.Scene(str(xml), TESTS, 24, 16, 4, 0, 0)
    d = orc.load_scene_description(str(xml))
    o = orc.Scene(d, dict(width=24, height=16, spp=4, sppe=0, sppse=0))
    for s in (r, o):
        s.set_envmap_tangent(None, 1.0)
    r.configure()
    o.configure()
    assert_images_close(refrun.DirectIntegrator(1, 1).renderC(r), orc.DirectIntegrator(1, 1).renderC(o), rel=2e-4, outliers=0.01, what="renderC")
    (a, at), (b, bt) = refrun.DirectIntegrator(1, 1).renderD(r), orc.DirectIntegrator(1, 1).renderD(o)
    assert np.abs(b).max() > 0 and np.abs(bt).max() > 0
    assert_images_close(a, b, rel=2e-4, outliers=0.01, what="renderD")
    assert_images_close(at, bt, rel=1e-3, outliers=0.02, what="tangent")


def test_reference_run_derivative_image_agrees_with_finite_differences_of_the_geometry():
    """the physical check of examples/run_test.py:150-231 on the reference run itself, independent of the oracle: translating the bunny of
    cbox_bunny along x, the forward-mode derivative image (interior + primary-edge + secondary-edge terms, carried by the stand-in's tangents
    through the reference's own statements) matches the central finite difference of the reference's renderC images; without the boundary
    terms it does not — what path-space differentiable rendering is about"""
    W, spp, eps = 32, 128, 1.0
    xml = scene_path("cbox_bunny")
    refrun.set_matvec_plain(True)

    def render(dx, terms):
        r = refrun.Scene(xml, TESTS, W, W, spp, spp if terms == "all" else 0, spp if terms == "all" else 0)
        M = np.eye(4, dtype=np.float32)
        M[0, 3] = dx
        r.set_mesh_transform(1, M, True)
        if terms:
            t = np.zeros((4, 4), np.float32)
            t[0, 3] = 1.0
            r.set_mesh_transform_tangent(1, t, True)
        r.configure()
        I = refrun.DirectIntegrator(1, 1)
        return I.renderD(r)[1] if terms else I.renderC(r)

    def blur(x):   # 4x4 box filter on the luminance-like channel mean: the estimators are noisy per pixel
        return x.reshape(W, W, 3).mean(2).reshape(W // 4, 4, W // 4, 4).mean((1, 3)).ravel()
    fd = blur((render(eps, None) - render(-eps, None)) / (2 * eps))
    full, interior = blur(render(0.0, "all")), blur(render(0.0, "interior"))
    corr = lambda a, b: float(np.corrcoef(a, b)[0, 1])
    assert corr(full, fd) > 0.85, corr(full, fd)
    assert corr(interior, fd) < 0.6, corr(interior, fd)
    assert 0.8 < np.linalg.norm(full) / np.linalg.norm(fd) < 1.25
