# vertex-position gradients of the interior term vs the oracle's forward mode (dot-product tests)
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import orc
from psdr_cuda_b200 import capi
desc = orc.load_scene_description('tests/data/scenes/cbox_bunny.xml')
W = H = 48; spp = 8
opts = dict(width=W, height=H, spp=spp, sppe=0, sppse=0)
rng = np.random.default_rng(11)
dLdI = rng.uniform(-1, 1, size=(W * H, 3)).astype(np.float32)
for kind, kw in (("direct", dict(bsdf_samples=1, light_samples=1)), ("path", dict(max_depth=3))):
    for mesh in (1, 2, 0, 5):
        ctx = capi.Context(0); ctx.load_description(desc, opts)
        ctx.grad_require(capi.PARAM_MESH_VERTICES, mesh)
        ctx.configure()
        pi = capi.make_integrator(kind, **kw)
        ctx.render_d(pi)
        t0 = time.time()
        g = ctx.render_d_vjp(pi, torch.from_numpy(dLdI).cuda()).cpu().numpy().reshape(-1, 3)
        dt = time.time() - t0
        oi = orc.DirectIntegrator(1, 1) if kind == "direct" else orc.PathIntegrator(kw['max_depth'])
        nv = len(desc['meshes'][mesh]['verts'])
        line = "%s mesh %d (nv=%d) |g|=%.4g vjp %.3fs:" % (kind, mesh, nv, np.linalg.norm(g), dt)
        for trial in range(3):
            if trial == 0: u = np.tile(np.array([[1.0, 0.5, -0.3]], np.float32), (nv, 1))      # rigid translation
            else: u = rng.normal(size=(nv, 3)).astype(np.float32)
            osc = orc.Scene(desc, opts)
            osc.set_mesh_vertex_tangent(mesh, u)
            osc.configure()
            _, dimg = oi.renderD(osc)
            want = float((dLdI.astype(np.float64) * dimg).sum())
            got = float((g.astype(np.float64) * u).sum())
            line += "  [%d] got %.5g want %.5g rel %.2e" % (trial, got, want, abs(got - want) / max(abs(want), 1e-9))
        print(line)
        ctx.close()
print("DONE")
