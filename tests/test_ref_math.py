"""Pins the oracle's restatement of psdr-cuda's per-lane math to the reference's OWN source: oracle/_ref/libref_math.so is
include/psdr/core/{warp,frame}.h, include/psdr/utils.h, include/psdr/bsdf/ggx.h and src/bsdf/ggx.cpp compiled unmodified from
/root/reference (oracle/build_ref.sh) against a scalar stand-in for Enoki (oracle/ref_stub). Covers SURVEY §8a rows a9 (warps), a10
(GGX distribution, visible-normal sampling, Smith G1, conductor Fresnel), a22 (frame, bilinear, sign, luminance), the ray/triangle
arithmetic that defines hit parity (utils.h:67-77) and the envmap's box exit (utils.h:129-145).

What stays assumed: Enoki's own semantics — exact 1/x, 1/sqrt(x) in place of its approximate rcp / rsqrt, libm
sin/cos/acos/atan2 — hence the tolerances of a few ulp where the reference normalises or takes a reciprocal, and the rendering path
as a whole (integrators, OptiX hits, autodiff), which cannot be built here."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import orc

REF = os.path.join(ROOT, "oracle", "_ref", "libref_math.so")
F = C.c_float


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libref_math.so not built (needs /root/reference: bash oracle/build_ref.sh)")
    ref, o = C.CDLL(REF), orc.lib()
    for L, pre in ((ref, "ref_"), (o, "orc_math_")):
        for name in ("rgb2luminance", "ggx_eval", "ggx_smith_g1"):
            getattr(L, pre + name).restype = C.c_float
    ref.ref_square_to_cosine_hemisphere_pdf.restype = C.c_float
    return ref, o


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def f32(*x):
    return np.array(x, np.float32).reshape(-1)


def both(libs, name, args, nout):
    """call ref_<name> and orc_math_<name> with the same arguments; the last argument is the output array"""
    out = []
    for L, pre in zip(libs, ("ref_", "orc_math_")):
        o = np.zeros(nout, np.float32)
        getattr(L, pre + name)(*args, p(o))
        out.append(o)
    return out


def close(a, b, ulps=4, atol=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= ulps * 1.2e-7 * np.maximum(np.abs(a), np.abs(b)) + atol)


def unit(rng):
    v = rng.normal(size=3)
    return (v / np.linalg.norm(v)).astype(np.float32)


def test_warps_match_reference_source(libs):
    rng = np.random.default_rng(1)
    for _ in range(2000):
        s = rng.uniform(0, 1, 2).astype(np.float32)
        for name, n in (("square_to_uniform_disk_concentric", 2), ("square_to_cosine_hemisphere", 3), ("square_to_uniform_triangle", 2)):
            r, o = both(libs, name, (p(s),), n)
            assert close(r, o, ulps=4, atol=2e-7), (name, s, r, o)


def test_frame_matches_reference_source(libs):
    rng = np.random.default_rng(2)
    for _ in range(2000):
        n, v = unit(rng), rng.normal(size=3).astype(np.float32)
        sr, tr = np.zeros(3, np.float32), np.zeros(3, np.float32); so, to = np.zeros(3, np.float32), np.zeros(3, np.float32)
        libs[0].ref_frame(p(n), p(sr), p(tr)); libs[1].orc_math_frame(p(n), p(so), p(to))
        assert np.array_equal(sr, so) and np.array_equal(tr, to), (n, sr, so)       # Duff et al. basis: no reciprocal approximations involved beyond 1/x
        for name in ("frame_to_local", "frame_to_world"):
            r, o = both(libs, name, (p(n), p(v)), 3)
            assert close(r, o, ulps=4, atol=3e-7), (name, r, o)


def test_ray_triangle_and_helpers_match_reference_source(libs):
    rng = np.random.default_rng(3)
    for _ in range(3000):
        p0, e1, e2 = (rng.normal(size=3).astype(np.float32) for _ in range(3))
        o, d = rng.normal(size=3).astype(np.float32) * 3, unit(rng)
        r, q = both(libs, "ray_intersect_triangle", (p(p0), p(e1), p(e2), p(o), p(d)), 3)
        # the reference multiplies by rcp(a): 1/a here, so the results agree to the last bit or two
        assert close(r, q, ulps=2, atol=1e-7 * max(1.0, float(np.abs(r).max()))), (r, q)
        st = rng.uniform(0, 1, 2).astype(np.float32)
        r, q = both(libs, "bilinear", (p(p0), p(e1), p(e2), p(st)), 3)
        assert np.array_equal(r, q)
        rgb = rng.uniform(0, 5, 3).astype(np.float32)
        assert close(libs[0].ref_rgb2luminance(p(rgb)), libs[1].orc_math_rgb2luminance(p(rgb)), ulps=2)
        x = np.float32(rng.normal() * 1e-5)
        assert libs[0].ref_sign_eps(F(x), F(1e-5)) == libs[1].orc_math_sign_eps(F(x), F(1e-5))
        lo, hi = f32(-2, -3, -1), f32(4, 2, 5)
        inside = (lo + (hi - lo) * rng.uniform(0.05, 0.95, 3)).astype(np.float32)
        r, q = both(libs, "ray_intersect_scene_aabb", (p(inside), p(d), p(lo), p(hi)), 5)
        assert np.array_equal(r[1:4], q[1:4]) and close(r[[0, 4]], q[[0, 4]], ulps=4), (r, q)


def test_fresnel_and_ggx_match_reference_source(libs):
    rng = np.random.default_rng(4)
    ref, o = libs
    for _ in range(3000):
        eta, k = rng.uniform(0.1, 2.0, 3).astype(np.float32), rng.uniform(0.5, 5.0, 3).astype(np.float32)
        c = F(rng.uniform(0.01, 1.0))
        r, q = both(libs, "fresnel", (p(eta), p(k), c), 3)
        assert close(r, q, ulps=8), (r, q)
        au, av = F(rng.uniform(0.03, 0.8)), F(rng.uniform(0.03, 0.8))
        m = unit(rng); m[2] = abs(m[2])
        wi = unit(rng); wi[2] = abs(wi[2]) + 1e-3; wi /= np.linalg.norm(wi)
        assert close(ref.ref_ggx_eval(au, av, p(m)), o.orc_math_ggx_eval(au, av, p(m)), ulps=8)
        assert close(ref.ref_ggx_smith_g1(au, av, p(wi), p(m)), o.orc_math_ggx_smith_g1(au, av, p(wi), p(m)), ulps=8)
        s2 = rng.uniform(0.01, 0.99, 2).astype(np.float32)
        r, q = both(libs, "ggx_sample_visible_11", (F(wi[2]), p(s2)), 2)
        assert close(r, q, ulps=16, atol=1e-6), (r, q)
        s3 = rng.uniform(0.01, 0.99, 3).astype(np.float32)
        r, q = both(libs, "ggx_sample", (au, av, p(wi), p(s3)), 3)
        # the stretched direction wi_p = normalize(alpha * wi.xy, wi.z) is nearly the normal for small alpha, and the reference then takes
        # sin = sqrt(1 - cos^2): a last-bit difference in cos (normalize = v * rsqrt(|v|^2) here, v / sqrt(|v|^2) in the oracle) is
        # amplified by 1 / sin^2
        wp = np.array([au.value * wi[0], av.value * wi[1], wi[2]], np.float64); wp /= np.linalg.norm(wp)
        amp = 1.0 / max(1e-6, wp[0] ** 2 + wp[1] ** 2)
        assert close(r, q, ulps=32, atol=2e-6 + 2.4e-7 * amp), (r, q, amp)


def test_bsdfs_match_reference_source(libs):
    """Diffuse and RoughConductor eval / pdf / sample (src/bsdf/diffuse.cpp, src/bsdf/roughconductor.cpp through src/core/bitmap.cpp with
    constant textures) against the oracle's Scene::bsdf_*: values, validity masks and which sample dimensions each BSDF consumes."""
    ref, o = libs
    ref.ref_diffuse_pdf.restype = ref.ref_rc_pdf.restype = o.orc_math_bsdf_pdf.restype = C.c_float
    rng = np.random.default_rng(6)
    n_valid = [0, 0]
    for it in range(3000):
        wi = unit(rng); wo = unit(rng)
        if it % 5:      # mostly the upper hemisphere, sometimes below the surface (masks)
            wi[2], wo[2] = abs(wi[2]), abs(wo[2])
        s3 = rng.uniform(0.01, 0.99, 3).astype(np.float32)
        # diffuse
        rho = rng.uniform(0, 1, 3).astype(np.float32)
        r, q = np.zeros(3, np.float32), np.zeros(3, np.float32)
        ref.ref_diffuse_eval(p(rho), p(wi), p(wo), p(r)); o.orc_math_bsdf_eval(0, p(rho), p(wi), p(wo), p(q))
        assert close(r, q, ulps=2), ("diffuse eval", r, q)
        assert close(ref.ref_diffuse_pdf(p(rho), p(wi), p(wo)), o.orc_math_bsdf_pdf(0, p(rho), p(wi), p(wo)), ulps=2)
        r4, q4 = np.zeros(4, np.float32), np.zeros(4, np.float32)
        vr = ref.ref_diffuse_sample(p(rho), p(wi), p(s3), p(r4)); vq = o.orc_math_bsdf_sample(0, p(rho), p(wi), p(s3), p(q4))
        assert vr == vq and close(r4, q4, ulps=4, atol=3e-7), ("diffuse sample", r4, q4)
        # rough conductor
        prm = np.concatenate([rng.uniform(0.08, 0.8, 2), rng.uniform(0.1, 2.0, 3), rng.uniform(0.5, 5.0, 3), rng.uniform(0.3, 1.0, 3)]).astype(np.float32)
        ref.ref_rc_eval(p(prm), p(wi), p(wo), p(r)); o.orc_math_bsdf_eval(1, p(prm), p(wi), p(wo), p(q))
        assert close(r, q, ulps=32, atol=1e-9), ("rc eval", prm, wi, wo, r, q)
        if wi[2] > 0:
            a, b = ref.ref_rc_pdf(p(prm), p(wi), p(wo)), o.orc_math_bsdf_pdf(1, p(prm), p(wi), p(wo))
            assert close(a, b, ulps=32, atol=1e-9), ("rc pdf", a, b)
            vr = ref.ref_rc_sample(p(prm), p(wi), p(s3), p(r4)); vq = o.orc_math_bsdf_sample(1, p(prm), p(wi), p(s3), p(q4))
            wp = np.array([prm[0] * wi[0], prm[1] * wi[1], wi[2]], np.float64); wp /= np.linalg.norm(wp)
            amp = 1.0 / max(1e-6, wp[0] ** 2 + wp[1] ** 2)
            assert vr == vq, ("rc sample validity", vr, vq, r4, q4)
            assert close(r4[:3], q4[:3], ulps=32, atol=2e-6 + 2.4e-7 * amp), ("rc sample wo", r4, q4)
            assert abs(r4[3] - q4[3]) <= (1e-4 + 3e-6 * amp) * max(abs(r4[3]), 1e-3), ("rc sample pdf", r4, q4)
            n_valid[vr] += 1
    assert n_valid[1] > 1000


def test_sampler_streams_match_reference_source(libs):
    """src/core/sampler.cpp (sample_tea_64 in 64-bit lanes with a 32-bit running sum, the seeding rule, next_1d) and sampler.h's next_2d /
    next_nd<3>, whose component order is the argument-evaluation order of the compiler — gcc here, as in the reference's documented
    Linux toolchain — against the oracle's SamplerLane, bit for bit."""
    ref, o = libs
    for lane in (0, 1, 2, 3, 12345, 67108863, 2**31 + 17, 2**33 + 5):
        outs = []
        for L, name in ((ref, "ref_sampler_lane"), (o, "orc_math_sampler_lane")):
            a, b, c = np.zeros(16, np.float32), np.zeros(2, np.float32), np.zeros(3, np.float32)
            getattr(L, name)(C.c_uint64(lane), 16, p(a), p(b), p(c))
            outs.append(np.concatenate([a, b, c]))
        assert np.array_equal(outs[0], outs[1]), (lane, outs)
    # and the known answers recorded in SURVEY §8c for streams 0-2
    a, b, c = np.zeros(3, np.float32), np.zeros(2, np.float32), np.zeros(3, np.float32)
    ref.ref_sampler_lane(C.c_uint64(0), 3, p(a), p(b), p(c))
    assert np.allclose(a, [0.79081202, 0.05460072, 0.82107019], atol=1e-7)


def test_product_device_math_matches_reference_source(tmp_path):
    """tests/native/ref_math_check.cu: the per-lane math the sm_100a kernels inline (csrc/pb_math.cuh: warps, frame, ray / triangle, bilinear,
    luminance, sampler streams incl. the jump-ahead; csrc/pb_rc.cuh: GGX, Smith G1, visible-normal sampling, conductor Fresnel, rough-conductor
    eval / pdf / sample), host-compiled, against the reference's own source in oracle/_ref/libref_math.so — function by function, no oracle
    in between"""
    import shutil
    import subprocess
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libref_math.so not built (needs /root/reference: bash oracle/build_ref.sh)")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "ref_math_check")
    subprocess.check_call([nvcc, "-x", "cu", "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "ref_math_check.cu"), "-ldl"], stderr=subprocess.DEVNULL)
    render = os.path.join(ROOT, "oracle", "_ref", "libref_render.so")   # adds the derivatives: the kernels' forward-mode duals of the rough conductor
    args = [exe, REF] + ([render] if os.path.exists(render) else [])    # vs the tangents of the reference's D flavour (roughconductor.cpp with its detach() calls)
    out = subprocess.run(args, capture_output=True, text=True)
    assert out.returncode == 0 and "ref_math_check: ok" in out.stdout, out.stdout[-3000:]


def test_product_shading_primitives_match_reference_source(tmp_path):
    """tests/native/ref_shade_check.cu: csrc/pb_shade.cuh as the primal kernels inline it — bitmap lookups, discrete sampling with sample reuse,
    the float GGX / Fresnel path, BSDF eval / pdf / sample for diffuse and rough-conductor records (both kernel instantiations), the scene-box
    exit of environment-map samples — host-compiled, against the reference's own ggx.cpp / diffuse.cpp / roughconductor.cpp / utils.h
    (libref_math.so) and bitmap.cpp / pmf.cpp (libref_render.so)"""
    import shutil
    import subprocess
    render = os.path.join(ROOT, "oracle", "_ref", "libref_render.so")
    if not (os.path.exists(REF) and os.path.exists(render)):
        pytest.skip("oracle/_ref libraries not built (needs /root/reference: bash oracle/build_ref.sh)")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "ref_shade_check")
    subprocess.check_call([nvcc, "-x", "cu", "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "ref_shade_check.cu"), "-ldl"], stderr=subprocess.DEVNULL)
    out = subprocess.run([exe, REF, render, os.path.join(ROOT, "tests")], capture_output=True, text=True)   # + camera rays and envmap lookups on the fixtures
    assert out.returncode == 0 and "ref_shade_check: ok" in out.stdout, out.stdout[-3000:]
