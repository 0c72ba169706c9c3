# roughconductor + envmap parity (bunny_env.xml, bunny_env_2.xml) and multi-emitter scene vs the oracle
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import numpy as np, torch
from oracle import orc
from psdr_cuda_b200 import capi
for scene, opts in (("bunny_env", dict(width=64, height=64, spp=8, sppe=0, sppse=0)), ("bunny_env_2", dict(width=64, height=36, spp=8, sppe=0, sppse=0)),
                    ("cbox_bunny_mutiemitter", dict(width=64, height=64, spp=8, sppe=0, sppse=0)), ("tree", dict(width=64, height=64, spp=8, sppe=0, sppse=0))):
    desc = orc.load_scene_description('tests/data/scenes/%s.xml' % scene)
    for kind, kw in (("direct", dict(bsdf_samples=1, light_samples=1)), ("direct", dict(bsdf_samples=2, light_samples=0)), ("direct", dict(bsdf_samples=0, light_samples=2)), ("path", dict(max_depth=3))):
        osc = orc.Scene(desc, opts); t0 = time.time(); osc.configure(); tc = time.time() - t0
        ctx = capi.Context(0); ctx.load_description(desc, opts); ctx.configure()
        oi = orc.DirectIntegrator(kw['bsdf_samples'], kw['light_samples']) if kind == "direct" else orc.PathIntegrator(kw['max_depth'])
        pi = capi.make_integrator(kind, **kw)
        ref = oi.renderC(osc); img = ctx.render_c(pi).cpu().numpy()
        refd, _ = oi.renderD(osc); imgd = ctx.render_d(pi).cpu().numpy()
        e, ed = np.abs(img - ref).mean(1), np.abs(imgd - refd).mean(1)
        ti_ok = np.array_equal(osc.triangle_info().view(np.uint32), ctx.triangle_info().view(np.uint32))
        print("%-24s %-6s %-40s tri-table %s | renderC mean %.5f/%.5f err max %.2e frac>1e-4 %.4f | renderD err max %.2e frac>1e-4 %.4f" %
              (scene, kind, kw, ti_ok, ref.mean(), img.mean(), e.max(), np.mean(e > 1e-4), ed.max(), np.mean(ed > 1e-4)), flush=True)
        ctx.close()
print("DONE")
