import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import scene_path
from oracle import orc
from psdr_cuda_b200 import capi, scene_io
scene = sys.argv[1] if len(sys.argv) > 1 else "cbox_bunny"
pdesc = scene_io.load_scene_description(scene_path(scene)); odesc = orc.load_scene_description(scene_path(scene))
RES = int(os.environ.get("RES", 32)); SPP = int(os.environ.get("SPP", 4))
opts = dict(width=RES, height=RES, spp=SPP, sppe=0, sppse=0)
rng = np.random.default_rng(1)
yy, xx = np.mgrid[0:RES, 0:RES]
dLdI = (1.0 + 0.5 * np.sin(xx / RES * 3.0 + 0.3)[..., None] * np.cos(yy / RES * 2.0)[..., None] * np.array([1.0, 0.8, 0.6])).reshape(-1, 3).astype(np.float32)
for label, kind, kw in (("both", "direct", dict(bsdf_samples=1, light_samples=1)),):
    ctx = capi.Context(0); ctx.load_description(pdesc, opts); ctx.grad_require(capi.PARAM_SENSOR_TRANSFORM, 0); ctx.configure()
    hide = len(sys.argv) > 3
    integ = capi.make_integrator(kind, hide_emitters=hide, **kw)
    img = ctx.render_d(integ).cpu().numpy()
    oi0 = orc.DirectIntegrator(kw["bsdf_samples"], kw["light_samples"], hide)
    osc0 = orc.Scene(odesc, opts); osc0.configure()
    ref_img, _ = oi0.renderD(osc0)
    err = np.abs(img - ref_img).mean(axis=1)
    print("image parity: px over 1e-4: %.4f, max err %.3g, mean %.5f vs %.5f" % (np.mean(err > 1e-4), err.max(), img.mean(), ref_img.mean()))
    g = ctx.render_d_vjp(integ, torch.from_numpy(dLdI).cuda()).cpu().numpy().reshape(4, 4)
    ref = np.zeros((4, 4))
    for i in range(3):
        for j in range(4):
            T = np.zeros((4, 4), np.float32); T[i, j] = 1
            osc = orc.Scene(odesc, opts); osc.set_sensor_transform_tangent(0, T); osc.configure()
            oi = orc.DirectIntegrator(kw["bsdf_samples"], kw["light_samples"], hide)
            _, dimg = oi.renderD(osc)
            ref[i, j] = float((dLdI.astype(np.float64) * dimg).sum())
    print(label); print(np.array2string(g[:3], precision=4)); print(np.array2string(ref[:3], precision=4)); print("diff", np.array2string(g[:3] - ref[:3], precision=3), flush=True)
    ctx.close()
