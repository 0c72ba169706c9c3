"""Generates tests/golden/ref_source_golden.npz: outputs of psdr-cuda's OWN renderer source, run on the CPU in this container through
oracle/_ref/libref_render.so (the reference's src/**/*.cpp compiled unmodified against stand-ins for Enoki and OptiX).
These vectors travel in git, so the oracle (CPU, tests/test_oracle.py) and the CUDA product (tests/test_gpu_parity.py, on the B200 box
where /root/reference does not exist) are both checked directly against what the reference's code computes.

    python tests/golden/make_ref_golden.py        (needs /root/reference; rebuilds the library if it is stale)

Each case: scene, RenderOption, DirectIntegrator(bsdf_samples, light_samples) or a field, optionally one forward-mode tangent;
stored: the renderC image of a freshly configured scene, or the renderD image + derivative image."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402

TESTS = os.path.join(ROOT, "tests")
TRANSLATE = [1.0, 0.5, -0.3]


def scene(name):
    return os.path.join(TESTS, "data", "scenes", name + ".xml")


def cases():
    """label -> (scene, (w, h, spp, sppe, sppse), integrator spec, leaf spec or None). Shared with the tests that replay them."""
    return {
        # BASELINE.json configs[0] at full size: DirectIntegrator renderC on cbox_bunny, 128x128, 16 spp
        "c_cfg1_full": ("cbox_bunny", (128, 128, 16, 0, 0), ("direct", 1, 1), None),
        "c_cbox_d11": ("cbox_bunny", (48, 48, 4, 0, 0), ("direct", 1, 1), None),
        "c_cbox_d21": ("cbox_bunny", (48, 48, 4, 0, 0), ("direct", 2, 1), None),
        "c_multi_d11": ("cbox_bunny_mutiemitter", (40, 40, 4, 0, 0), ("direct", 1, 1), None),
        "c_env_d11": ("bunny_env", (40, 40, 4, 0, 0), ("direct", 1, 1), None),
        "c_env2_d02": ("bunny_env_2", (48, 27, 4, 0, 0), ("direct", 0, 2), None),
        "c_cbox_position": ("cbox_bunny", (48, 48, 1, 0, 0), ("field", "position"), None),
        "c_cbox_shnormal": ("cbox_bunny", (48, 48, 1, 0, 0), ("field", "shNormal"), None),
        "d_cbox_albedo": ("cbox_bunny", (48, 48, 4, 0, 0), ("direct", 1, 1), ("bsdf", 0, "reflectance", [1.0, 0.5, 0.25])),
        "d_cbox_translate": ("cbox_bunny", (48, 48, 4, 0, 0), ("direct", 1, 1), ("translate", 1)),
        # 3 edge samples per pixel: an edge-ray pair that straddles a silhouette by 1e-5 is all-or-nothing under last-bit differences; fewer lanes
        # per silhouette pixel keep the share of pixels that contain such a lane at a few per cent (8 samples per pixel: 4.5 %)
        "d_cbox_primary": ("cbox_bunny", (48, 48, 0, 3, 0), ("direct", 1, 1), ("translate", 1)),
        "d_cbox_secondary": ("cbox_bunny", (48, 48, 0, 0, 32), ("direct", 1, 1), ("translate", 1)),
        "d_env_alpha": ("bunny_env", (32, 32, 4, 0, 0), ("direct", 1, 1), ("bsdf", 0, "alpha_u", [1.0])),
        "d_env_scale": ("bunny_env", (32, 32, 4, 0, 0), ("direct", 1, 1), ("env_scale",)),
        "d_env_translate": ("bunny_env", (32, 32, 4, 0, 0), ("direct", 1, 1), ("translate", 0)),
        # one BSDF + environment map: boundary segments ending on the bounding mesh are shaded with meshes[0]'s BSDF (direct.cpp:278-284)
        "d_env_secondary": ("bunny_env", (32, 32, 0, 0, 16), ("direct", 1, 1), ("translate", 0)),
    }


def seed(sc, leaf, num_vertices):
    """sets the case's tangent on a refrun.Scene or an orc.Scene (same method names)"""
    if leaf is None:
        return
    if leaf[0] == "bsdf":
        sc.set_bsdf_tangent(leaf[1], leaf[2], np.asarray(leaf[3], np.float32).reshape(1, 1, -1))
    elif leaf[0] == "translate":
        sc.set_mesh_vertex_tangent(leaf[1], np.tile(np.asarray([TRANSLATE], np.float32), (num_vertices, 1)))
    elif leaf[0] == "env_scale":
        sc.set_envmap_tangent(None, 1.0)
    else:
        raise ValueError(leaf)


def main():
    refrun.set_matvec_plain(True)   # matrix * vector in the oracle's / product's form (plain sums): see oracle/ref_dyn/enoki_dyn.h
    out = {}
    for label, (name, (w, h, spp, sppe, sppse), integ, leaf) in cases().items():
        sc = refrun.Scene(scene(name), TESTS, w, h, spp, sppe, sppse)
        seed(sc, leaf, sc.num_vertices(leaf[1]) if leaf and leaf[0] == "translate" else 0)
        sc.configure()
        I = refrun.DirectIntegrator(integ[1], integ[2]) if integ[0] == "direct" else refrun.FieldExtractionIntegrator(integ[1])
        if leaf is None:
            out[label] = I.renderC(sc)
        else:
            img, dimg = I.renderD(sc)
            out[label], out[label + "_t"] = img, dimg
        print(label, "max %.4g" % np.abs(out[label]).max(), "" if leaf is None else "tangent max %.4g" % np.abs(out[label + "_t"]).max())
    # a slice of the configure tables of cbox_bunny (every 97th row) + whole-table sums
    sc = refrun.Scene(scene("cbox_bunny"), TESTS, 16, 16, 1, 1, 1)
    sc.configure()
    tri, sec, prim = sc.triangle_info(), sc.sec_edges(), sc.primary_edges()
    out["t_tri_rows"], out["t_sec_rows"], out["t_prim_rows"] = tri[::97], sec[::97], prim[::97]
    out["t_counts"] = np.array([len(tri), len(sec), len(prim)], np.int64)
    out["t_tri_sum"], out["t_sec_sum"] = tri.astype(np.float64).sum(0), sec.astype(np.float64).sum(0)
    path = os.path.join(TESTS, "golden", "ref_source_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
