import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import orc
from psdr_cuda_b200 import capi

desc = orc.load_scene_description('tests/data/scenes/cbox_bunny.xml')
opts = dict(width=64, height=64, spp=8, sppe=0, sppse=0)
rng = np.random.default_rng(7)
dLdI = rng.uniform(-1, 1, size=(64 * 64, 3)).astype(np.float32)
for kind, kw in (("direct", dict(bsdf_samples=1, light_samples=1)), ("direct", dict(bsdf_samples=2, light_samples=1)), ("path", dict(max_depth=1)), ("path", dict(max_depth=4))):
    ctx = capi.Context(0); ctx.load_description(desc, opts)
    for b in range(4):
        ctx.grad_require(capi.PARAM_BSDF_TEXTURE, b, "reflectance")
    ctx.configure()
    pi = capi.make_integrator(kind, **kw)
    imgd = ctx.render_d(pi).cpu().numpy()
    grad = ctx.render_d_vjp(pi, torch.from_numpy(dLdI).cuda()).cpu().numpy()
    # oracle: renderD primal + JVPs
    oi = orc.DirectIntegrator(kw['bsdf_samples'], kw['light_samples']) if kind == "direct" else orc.PathIntegrator(kw['max_depth'])
    ref_g = np.zeros(12, np.float32)
    for b in range(4):
        for ch in range(3):
            osc = orc.Scene(desc, opts)
            t = np.zeros((1, 1, 3), np.float32); t[0, 0, ch] = 1
            osc.set_bsdf_tangent(b, "reflectance", t)
            osc.configure()
            img_o, dimg = oi.renderD(osc)
            ref_g[3 * b + ch] = float((dLdI.astype(np.float64) * dimg).sum())
    d = np.abs(imgd - img_o).mean(axis=1)
    print(kind, kw, "renderD primal: max %.3e mean %.3e" % (d.max(), d.mean()))
    print("  grad gpu", grad[:12])
    print("  grad ref", ref_g)
    print("  rel L2 err %.3e" % (np.linalg.norm(grad[:12] - ref_g) / np.linalg.norm(ref_g)))
    ctx.close()
print("DONE")
