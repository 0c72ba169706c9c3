"""Regenerate tests/golden/derivative_golden.npz from the CPU oracle (`python tests/golden/make_golden_derivatives.py`).

Pins the oracle's forward-mode derivative images (projected onto fixed weights) for every kind of leaf: rough-conductor parameters,
environment-map radiance / scale / transform, mesh vertices / texture coordinates, sensor pose — interior and boundary terms. Like
cbox_bunny_golden.npz these fixtures pin the oracle against itself; parity against the real psdr-cuda remains unpinned.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def scene(name):
    return os.path.join(ROOT, "tests", "data", "scenes", name + ".xml")


def weights(n):
    return np.random.default_rng(99).uniform(-1, 1, size=(n, 3)).astype(np.float32)


def cases():
    """(label, scene, opts, integrator factory, function that sets the tangent on a fresh oracle scene)"""
    o = dict(width=24, height=24, spp=4, sppe=0, sppse=0)
    oe = dict(width=24, height=24, spp=4, sppe=4, sppse=4)
    d11 = lambda: orc.DirectIntegrator(1, 1)
    p2 = lambda: orc.PathIntegrator(2)
    rng = np.random.default_rng(5)
    env_t = rng.normal(size=(512, 1024, 3)).astype(np.float32)
    out = [
        ("rc_alpha_u", "bunny_env", o, d11, lambda s: s.set_bsdf_tangent(0, "alpha_u", np.ones((1, 1, 1), np.float32))),
        ("rc_eta_path2", "bunny_env", o, p2, lambda s: s.set_bsdf_tangent(0, "eta", np.array([[[1.0, -0.5, 0.25]]], np.float32))),
        ("rc_k", "cbox_bunny_rc", o, p2, lambda s: s.set_bsdf_tangent(3, "k", np.array([[[0.3, 1.0, -0.7]]], np.float32))),
        ("env_scale", "bunny_env_2", o, d11, lambda s: s.set_envmap_tangent(None, 1.0)),
        ("env_radiance", "bunny_env_2", o, p2, lambda s: s.set_envmap_tangent(env_t, 0.0)),
        ("env_transform", "bunny_env_2", o, d11, lambda s: s.set_envmap_transform_tangent(np.array([[0, -1, 0, 0], [1, 0, 0.5, 0], [0, -0.5, 0, 0], [0, 0, 0, 0]], np.float32))),
        ("sensor_translate_all_terms", "cbox_bunny", oe, d11, lambda s: s.set_sensor_transform_tangent(0, np.array([[0, 0, 0, 1], [0, 0, 0, 0.5], [0, 0, 0, 0], [0, 0, 0, 0]], np.float32))),
        ("sensor_rotate_rc", "cbox_bunny_rc", o, p2, lambda s: s.set_sensor_transform_tangent(0, np.array([[0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]], np.float32))),
        ("vertices_rc_bunny", "cbox_bunny_rc", o, p2, lambda s: s.set_mesh_vertex_tangent(1, np.tile(np.array([[1.0, 0.5, -0.3]], np.float32), (34817, 1)))),
        ("vertices_env_floor", "bunny_env_2", o, d11, lambda s: s.set_mesh_vertex_tangent(1, np.tile(np.array([[0.2, -0.4, 1.0]], np.float32), (4, 1)))),
    ]
    return out


def compute():
    res = {}
    for label, sc_name, opts, make_integ, set_tangent in cases():
        desc = orc.load_scene_description(scene(sc_name))
        s = orc.Scene(desc, opts)
        set_tangent(s)
        s.configure()
        img, dimg = make_integ().renderD(s)
        w = weights(img.shape[0]).astype(np.float64)
        res[label] = np.array([float((w * dimg).sum()), float(np.abs(dimg).sum()), float(img.astype(np.float64).sum())])
    return res


if __name__ == "__main__":
    r = compute()
    np.savez(os.path.join(OUT, "derivative_golden.npz"), **r)
    for k, v in r.items():
        print(k, v)
