# boundary terms (primary / secondary edges) vs the oracle's forward mode (dot-product tests), one term at a time
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import orc
from psdr_cuda_b200 import capi
which = sys.argv[1] if len(sys.argv) > 1 else "all"
rng = np.random.default_rng(11)
def run(scene, opts, kind, kw, mesh, label, guide=None):
    desc = orc.load_scene_description('tests/data/scenes/%s.xml' % scene)
    W, H = opts['width'], opts['height']
    dLdI = rng.uniform(-1, 1, size=(W * H, 3)).astype(np.float32)
    ctx = capi.Context(0); ctx.load_description(desc, opts)
    ctx.grad_require(capi.PARAM_MESH_VERTICES, mesh)
    ctx.configure()
    pi = capi.make_integrator(kind, use_guiding=guide is not None, **kw)
    if kind == "direct": oi = orc.DirectIntegrator(kw.get('bsdf_samples', 1), kw.get('light_samples', 1))
    elif kind == "path": oi = orc.PathIntegrator(kw['max_depth'])
    else: oi = orc.FieldExtractionIntegrator(kw['field'])
    if guide is not None:
        t0 = time.time(); ctx.preprocess_secondary_edges(0, guide[0], guide[1]); tg = time.time() - t0
    ctx.render_d(pi)
    t0 = time.time()
    g = ctx.render_d_vjp(pi, torch.from_numpy(dLdI).cuda()).cpu().numpy().reshape(-1, 3)
    dt = time.time() - t0
    nv = len(desc['meshes'][mesh]['verts'])
    line = "%s %s %s mesh %d |g|=%.4g vjp %.3fs:" % (label, scene, kind, mesh, np.linalg.norm(g), dt)
    for trial in range(3):
        if trial == 0: u = np.tile(np.array([[1.0, 0.5, -0.3]], np.float32), (nv, 1))
        else: u = rng.normal(size=(nv, 3)).astype(np.float32)
        osc = orc.Scene(desc, opts)
        osc.set_mesh_vertex_tangent(mesh, u)
        osc.configure()
        if guide is not None: oi.preprocess_secondary_edges(osc, 0, guide[0], guide[1])
        _, dimg = oi.renderD(osc)
        want = float((dLdI.astype(np.float64) * dimg).sum())
        got = float((g.astype(np.float64) * u).sum())
        line += "  [%d] got %.5g want %.5g rel %.2e" % (trial, got, want, abs(got - want) / max(abs(want), 1e-9))
    print(line, flush=True)
    ctx.close()
if which in ("all", "prim"):
    run("bunny", dict(width=64, height=64, spp=0, sppe=16, sppse=0), "field", dict(field="silhouette"), 0, "primary-only(field)")
    run("cbox_bunny", dict(width=48, height=48, spp=0, sppe=8, sppse=0), "direct", dict(bsdf_samples=1, light_samples=1), 1, "primary-only")
    run("cbox_bunny", dict(width=48, height=48, spp=0, sppe=8, sppse=0), "path", dict(max_depth=2), 1, "primary-only")
if which in ("all", "sec"):
    run("cbox_bunny", dict(width=48, height=48, spp=0, sppe=0, sppse=32), "direct", dict(bsdf_samples=1, light_samples=1), 1, "secondary-only")
    run("cbox_bunny", dict(width=48, height=48, spp=0, sppe=0, sppse=32), "direct", dict(bsdf_samples=1, light_samples=1), 0, "secondary-only(emitter)")
    run("cbox_bunny", dict(width=48, height=48, spp=0, sppe=0, sppse=32), "direct", dict(bsdf_samples=1, light_samples=1), 2, "secondary-only(floor)")
if which in ("all", "guide"):
    run("cbox_bunny", dict(width=48, height=48, spp=0, sppe=0, sppse=16), "direct", dict(bsdf_samples=1, light_samples=1), 1, "secondary-guided", guide=([200, 4, 4, 2], 2))
if which in ("all", "full"):
    run("cbox_bunny", dict(width=48, height=48, spp=8, sppe=8, sppse=8), "direct", dict(bsdf_samples=1, light_samples=1), 1, "all-terms")
    run("cbox_bunny", dict(width=48, height=48, spp=8, sppe=8, sppse=8), "path", dict(max_depth=3), 1, "all-terms")
print("DONE")
