"""The call sequence of the reference's examples (examples/run_test.py run_orig / run_ad / run_fd with the helpers of
examples/utils/differential.py), written against the same names: `import psdr_cuda`, `import enoki as ek`,
`enoki.cuda_autodiff.{Float32, Vector3f, Matrix4f}`. Run as a script by tests/test_examples_flow.py; prints a JSON summary."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import psdr_cuda_b200.compat  # noqa: E402,F401  (puts the psdr_cuda / enoki stand-ins on sys.path)

import psdr_cuda  # noqa: E402
import enoki as ek  # noqa: E402
from enoki.cuda_autodiff import Float32 as FloatD, Matrix4f as Matrix4fD, Vector3f as Vector3fD  # noqa: E402

scene_file = os.path.join(ROOT, "tests", "data", "scenes", sys.argv[1])
kind = sys.argv[2]


def mesh_transform(sc, mesh_ID, dir_vector):
    sc.param_map["Mesh[" + str(mesh_ID) + "]"].set_transform(Matrix4fD.translate(dir_vector))


def mesh_rotate(sc, mesh_ID, axis, angle):
    sc.param_map["Mesh[" + str(mesh_ID) + "]"].set_transform(Matrix4fD.rotate(axis, angle))


def vertex_transform(sc, mesh_ID, vertex_ID, dir_vector, orig_vtx_pos, P):
    para = "Mesh[" + str(mesh_ID) + "]"
    n = sc.param_map[para].num_vertices
    x_vals, y_vals, z_vals = [0.] * n, [0.] * n, [0.] * n
    x_vals[vertex_ID], y_vals[vertex_ID], z_vals[vertex_ID] = dir_vector
    sc.param_map[para].vertex_positions = Vector3fD(orig_vtx_pos) + Vector3fD(x_vals, y_vals, z_vals) * P


def apply(sc, P, orig=None):
    if kind == "mesh_transform":
        mesh_transform(sc, 1, Vector3fD([1.0, 0.0, 0.0]) * P)
    elif kind == "mesh_rotate":
        mesh_rotate(sc, 0, Vector3fD([0., 0.1, 0.]), P)
        mesh_rotate(sc, 1, Vector3fD([0., -0.1, 0.]), P)
    else:
        vertex_transform(sc, 0, 0, [-50.0, 0.0, 0.0], orig, P)


field = kind == "mesh_rotate"
make_integrator = (lambda: psdr_cuda.FieldExtractionIntegrator("silhouette")) if field else (lambda: psdr_cuda.DirectIntegrator(bsdf_samples=2, light_samples=2))

# ---- run_ad
sc = psdr_cuda.Scene()
sc.load_file(scene_file, False)
sc.opts.log_level = 0
sc.opts.width, sc.opts.height = 64, 64
sc.opts.spp, sc.opts.sppe, sc.opts.sppse = (16, 16, 0) if field else (8, 8, 16)
integrator = make_integrator()
orig = ek.detach(sc.param_map["Mesh[0]"].vertex_positions) if kind == "vertex_transform" else None
npass, img_ad = 4, None
for i in range(npass):
    P = FloatD(0.)
    ek.set_requires_gradient(P)
    apply(sc, P, orig)
    sc.configure()
    if i == 0 and not field:
        integrator.preprocess_secondary_edges(sc, 0, np.array([2000, 4, 4, 2]), 2)
    img = integrator.renderD(sc, 0)
    ek.forward(P, free_graph=True)
    grad_img = ek.gradient(img).numpy()
    grad_img[np.logical_not(np.isfinite(grad_img))] = 0.
    img_ad = grad_img if i == 0 else img_ad + grad_img
    del img, P
img_ad = (img_ad / float(npass)).reshape((sc.opts.height, sc.opts.width, 3))
del sc, integrator

# ---- run_fd
eps = 0.05 if kind != "vertex_transform" else 0.02
sc1, sc2 = psdr_cuda.Scene(), psdr_cuda.Scene()
for s in (sc1, sc2):
    s.load_file(scene_file, False)
    s.opts.width, s.opts.height, s.opts.spp = 64, 64, 64
    s.opts.sppe, s.opts.sppse, s.opts.log_level = 0, 0, 0
orig1 = ek.detach(sc1.param_map["Mesh[0]"].vertex_positions) if kind == "vertex_transform" else None
apply(sc1, FloatD(-eps), orig1)
apply(sc2, FloatD(eps), orig1)
sc1.configure()
sc2.configure()
integrator = make_integrator()
npass_fd, i1, i2 = 8, None, None
for i in range(npass_fd):
    a, b = integrator.renderC(sc1).numpy(), integrator.renderC(sc2).numpy()
    i1, i2 = (a, b) if i == 0 else (i1 + a, i2 + b)
img_fd = ((i2 - i1) / (2.0 * eps * float(npass_fd))).reshape((64, 64, 3))


def blocks(x, b=16):
    return x.reshape(64 // b, b, 64 // b, b, 3).mean(axis=(1, 3, 4))


A, F = blocks(img_ad), blocks(img_fd)
print(json.dumps({"sum_ad": float(img_ad.sum()), "sum_fd": float(img_fd.sum()), "corr": float(np.corrcoef(A.ravel(), F.ravel())[0, 1]),
                  "finite": bool(np.isfinite(img_ad).all()), "nonzero": float(np.abs(img_ad).max())}))
