"""Regenerate tests/golden/*.npz from the CPU oracle (run in the build container: `python tests/golden/make_golden.py`).

The reference ships no golden vectors, KATs or assertions and cannot be built here, so these
fixtures pin the oracle against itself (regression) and carry the canonical PCG32 known-answer vector; parity against
the real psdr-cuda remains unpinned. Everything here is produced by oracle/ only.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SCENE = os.path.join(ROOT, "tests", "data", "scenes", "cbox_bunny.xml")


def main():
    desc = orc.load_scene_description(SCENE)
    out = {}
    # RNG
    out["pcg32_42_54"] = orc.pcg32_kat(42, 54, 6)
    for lane in (0, 1, 2, 12345, 67108863):
        u, f = orc.sampler_kat(lane, 8)
        out["sampler_u32_%d" % lane] = u
        out["sampler_f32_%d" % lane] = f
    # cfg1-style render, small
    opts = dict(width=32, height=32, spp=4, sppe=0, sppse=0)
    for name, integ in (("direct11", orc.DirectIntegrator(1, 1)), ("direct21", orc.DirectIntegrator(2, 1)), ("path3", orc.PathIntegrator(3)),
                        ("field_depth", orc.FieldExtractionIntegrator("depth")), ("field_shn", orc.FieldExtractionIntegrator("shNormal"))):
        sc = orc.Scene(desc, opts)
        sc.configure()
        out["renderC_" + name] = integ.renderC(sc)
        out["renderC2_" + name] = integ.renderC(sc)   # second call continues the sampler streams
    # renderD + forward-mode derivative w.r.t. the white albedo (all three channels)
    sc = orc.Scene(desc, opts)
    sc.set_bsdf_tangent(0, "reflectance", np.ones((1, 1, 3), np.float32))
    sc.configure()
    img, dimg = orc.PathIntegrator(3).renderD(sc)
    out["renderD_path3"] = img
    out["renderD_path3_dwhite"] = dimg
    # fixed ray set and its hits
    rng = np.random.default_rng(2024)
    n = 4096
    o = np.stack([rng.uniform(-90, 90, n), rng.uniform(5, 190, n), rng.uniform(-90, 190, n)], 1).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    tri, shape, u, v, t = sc.trace(o, d)
    tri_b, shape_b, u_b, v_b, t_b = sc.trace(o, d, brute=True)
    assert np.array_equal(tri, tri_b) and np.array_equal(u, u_b) and np.array_equal(t, t_b), "oracle BVH != brute force"
    out.update(trace_o=o, trace_d=d, trace_tri=tri, trace_shape=shape, trace_u=u, trace_v=v, trace_t=t)
    # triangle table checksums
    ti = sc.triangle_info()
    out["tri_info_sum"] = ti.astype(np.float64).sum(axis=0)
    out["tri_info_first"] = ti[:8]
    out["tri_info_last"] = ti[-8:]
    out["edges_bunny_head"] = sc.mesh_edges(1)[:64]
    out["edges_counts"] = np.array([len(sc.mesh_edges(m)) for m in range(len(desc["meshes"]))])
    np.savez_compressed(os.path.join(OUT, "cbox_bunny_golden.npz"), **out)
    print("wrote", os.path.join(OUT, "cbox_bunny_golden.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
