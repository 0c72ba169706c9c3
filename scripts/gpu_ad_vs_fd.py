# The reference's own validation methodology (examples/run_test.py:44-231): forward-mode derivative image of renderD vs a
# central finite difference of renderC over perturbed scenes, here for a translation of the bunny in cbox_bunny.xml.
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psdr_cuda_b200 import capi, scene_io
W = H = 128
desc = scene_io.load_scene_description('tests/data/scenes/cbox_bunny.xml')
nv = len(desc['meshes'][1]['verts'])
axis = np.array([1.0, 0.0, 0.0], np.float32)
# the tangent is a world-space translation; vertex_positions are object space: u_obj = M^-1(3x3) axis
M = desc['meshes'][1]['to_world'][:3, :3].astype(np.float64)
u_obj = np.linalg.solve(M, axis.astype(np.float64)).astype(np.float32)
def make(spp, sppe, sppse, shift=0.0):
    ctx = capi.Context(0)
    ctx.load_description(desc, dict(width=W, height=H, spp=spp, sppe=sppe, sppse=sppse))
    ctx.grad_require(capi.PARAM_MESH_VERTICES, 1)
    if shift != 0.0:
        T = np.eye(4, dtype=np.float32); T[:3, 3] = axis * shift
        ctx.set_mesh_transform(1, T, True)
    ctx.configure()
    return ctx
for kind, kw in (("direct", dict(bsdf_samples=1, light_samples=1)), ("path", dict(max_depth=3))):
    integ = capi.make_integrator(kind, **kw)
    npass, spp = 16, 64
    ctx = make(spp, spp, spp)
    tang = torch.from_numpy(np.tile(u_obj[None, :], (nv, 1)).reshape(-1)).cuda()
    ad = torch.zeros((W * H, 3), device="cuda"); ad_int = torch.zeros_like(ad)
    t0 = time.time()
    for p in range(npass):
        ctx.render_d(integ)
        ad += ctx.render_d_jvp(integ, tang)
    ad /= npass
    t_ad = time.time() - t0
    ctx.close()
    # interior-only derivative (what you get without the boundary terms)
    ctx = make(spp, 0, 0)
    for p in range(npass):
        ctx.render_d(integ); ad_int += ctx.render_d_jvp(integ, tang)
    ad_int /= npass
    ctx.close()
    eps = 0.5
    fd = torch.zeros_like(ad)
    cp, cm = make(4 * spp, 0, 0, +eps), make(4 * spp, 0, 0, -eps)
    for p in range(npass):
        fd += (cp.render_c(integ) - cm.render_c(integ)) / (2 * eps)
    fd /= npass
    cp.close(); cm.close()
    def blocks(x, b=16):
        return x.view(H // b, b, W // b, b, 3).mean(dim=(1, 3, 4)).cpu().numpy()
    A, F, AI = blocks(ad), blocks(fd), blocks(ad_int)
    corr = lambda a, b: float(np.corrcoef(a.ravel(), b.ravel())[0, 1])
    print("%s %s: sum AD %.4f  sum FD %.4f  sum AD(interior only) %.4f | block corr AD~FD %.4f  interior-only~FD %.4f | rel block L2 AD %.3f interior-only %.3f  (AD %.1fs)" %
          (kind, kw, float(ad.sum()), float(fd.sum()), float(ad_int.sum()), corr(A, F), corr(AI, F),
           np.linalg.norm(A - F) / np.linalg.norm(F), np.linalg.norm(AI - F) / np.linalg.norm(F), t_ad), flush=True)
    np.save('gpurun_out/adfd_%s_ad.npy' % kind, ad.cpu().numpy()); np.save('gpurun_out/adfd_%s_fd.npy' % kind, fd.cpu().numpy())
print("DONE")
