# forward mode (JVP) derivative images vs the oracle's forward-mode duals, pixel by pixel
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import orc
from psdr_cuda_b200 import capi
desc = orc.load_scene_description('tests/data/scenes/cbox_bunny.xml')
rng = np.random.default_rng(5)
def run(opts, kind, kw, what, label):
    ctx = capi.Context(0); ctx.load_description(desc, opts)
    if what == "albedo": ctx.grad_require(capi.PARAM_BSDF_TEXTURE, 0, "reflectance")
    else: ctx.grad_require(capi.PARAM_MESH_VERTICES, 1)
    ctx.configure()
    pi = capi.make_integrator(kind, **kw)
    oi = orc.DirectIntegrator(kw.get('bsdf_samples', 1), kw.get('light_samples', 1)) if kind == "direct" else orc.PathIntegrator(kw['max_depth'])
    nv = len(desc['meshes'][1]['verts'])
    if what == "albedo": u = np.array([1.0, 0.5, 0.25], np.float32)
    elif what == "translate": u = np.tile(np.array([[1.0, 0.5, -0.3]], np.float32), (nv, 1))
    else: u = rng.normal(size=(nv, 3)).astype(np.float32)
    ctx.render_d(pi)
    t0 = time.time(); dimg = ctx.render_d_jvp(pi, torch.from_numpy(u.reshape(-1)).cuda()).cpu().numpy(); dt = time.time() - t0
    osc = orc.Scene(desc, opts)
    if what == "albedo": osc.set_bsdf_tangent(0, "reflectance", u.reshape(1, 1, 3))
    else: osc.set_mesh_vertex_tangent(1, u)
    osc.configure()
    _, ref = oi.renderD(osc)
    err = np.abs(dimg - ref)
    print("%-34s %s %s |ref| mean %.4g  abs err mean %.3e max %.3e  rel L2 %.3e  (jvp %.3fs)" % (label, kind, what, np.abs(ref).mean(), err.mean(), err.max(), np.linalg.norm(dimg - ref) / np.linalg.norm(ref), dt), flush=True)
    # JVP / VJP consistency: <v, J u> == <J^T v, u>
    v = rng.uniform(-1, 1, size=dimg.shape).astype(np.float32)
    g = ctx.render_d_vjp(pi, torch.from_numpy(v).cuda()).cpu().numpy()
    a, b = float((v.astype(np.float64) * dimg).sum()), float((g.astype(np.float64) * u.reshape(-1)).sum())
    print("    <v,Ju> %.6g  <J^T v,u> %.6g  rel %.2e" % (a, b, abs(a - b) / max(abs(a), 1e-12)))
    ctx.close()
o = dict(width=48, height=48, spp=8, sppe=0, sppse=0)
run(o, "path", dict(max_depth=3), "albedo", "interior")
run(o, "direct", dict(bsdf_samples=1, light_samples=1), "translate", "interior")
run(o, "path", dict(max_depth=3), "random", "interior")
run(dict(width=48, height=48, spp=0, sppe=8, sppse=0), "direct", dict(bsdf_samples=1, light_samples=1), "translate", "primary edges only")
run(dict(width=48, height=48, spp=0, sppe=0, sppse=32), "direct", dict(bsdf_samples=1, light_samples=1), "translate", "secondary edges only")
run(dict(width=48, height=48, spp=8, sppe=8, sppse=8), "path", dict(max_depth=2), "translate", "all terms")
print("DONE")
