import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import orc
from psdr_cuda_b200 import capi
desc = orc.load_scene_description('tests/data/scenes/cbox_bunny.xml')
W = H = 32; spp = 4
opts = dict(width=W, height=H, spp=spp, sppe=0, sppse=0)
ctx = capi.Context(0); ctx.load_description(desc, opts)
ctx.grad_require(capi.PARAM_BSDF_TEXTURE, 0, "reflectance")
ctx.configure()
for depth in (3,):
    pi = capi.make_integrator("path", max_depth=depth)
    img = ctx.render_d(pi).cpu().numpy()
    ptr, nbytes = C.c_void_p(), C.c_int64()
    capi.lib().pb_debug_retained_rad(ctx.h, C.byref(ptr), C.byref(nbytes))
    n = W * H * spp
    rad = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(rad.data_ptr()), ptr, C.c_size_t(n * 16), 3)
    rad = rad.cpu().numpy()[:, :3]
    osc = orc.Scene(desc, opts); osc.configure()
    oi = orc.PathIntegrator(depth)
    ref, _ = oi.renderD(osc)
    err = np.abs(img - ref).mean(axis=1)
    bad = np.argwhere(err > 1e-4).ravel()
    print("bad pixels", bad, err[bad])
    L = orc.lib()
    for pix in bad:
        for s in range(spp):
            lane = pix * spp + s
            out = np.zeros(3, np.float32)
            L.orc_debug_lane(osc.h, oi.h, 0, C.c_int64(int(lane)), 1, out.ctypes.data_as(C.c_void_p))
            outc = np.zeros(3, np.float32)
            L.orc_debug_lane(osc.h, oi.h, 0, C.c_int64(int(lane)), 0, outc.ctypes.data_as(C.c_void_p))
            print(" lane", lane, "gpu D", rad[lane], "oracle D", out, "oracle C", outc)
for b in range(1, 4): ctx.grad_require(capi.PARAM_BSDF_TEXTURE, b, "reflectance")
ctx.configure(reseed=True)
pi = capi.make_integrator("path", max_depth=3)
ctx.render_d(pi)
rng = np.random.default_rng(12345)
dLdI = rng.uniform(-1, 1, size=(32 * 32, 3)).astype(np.float32)
g = ctx.render_d_vjp(pi, torch.from_numpy(dLdI).cuda()).cpu().numpy()
ref = np.zeros(12)
for b in range(4):
    for ch in range(3):
        osc = orc.Scene(desc, opts)
        t = np.zeros((1, 1, 3), np.float32); t[0, 0, ch] = 1
        osc.set_bsdf_tangent(b, "reflectance", t)
        osc.configure()
        ref[3 * b + ch] = float((dLdI.astype(np.float64) * orc.PathIntegrator(3).renderD(osc)[1]).sum())
print("g  ", g)
print("ref", ref)
print("rel", np.linalg.norm(g - ref) / np.linalg.norm(ref), np.abs(g - ref))
