# BASELINE.json configs[4], one GPU's share: bunny_env.xml (rough conductor alpha 0.05 + environment map) at 1024x1024 with
# 64 spp (= 512 spp / 8 GPUs), texture (alpha_u, alpha_v, eta, k) + envmap scale + bunny vertex gradients, interior term and
# with both boundary terms. Prints timings per stage.
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psdr_cuda_b200 import capi, scene_io
W = H = int(os.environ.get("CFG5_RES", 1024)); SPP = int(os.environ.get("CFG5_SPP", 64))
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 1
desc = scene_io.load_scene_description('tests/data/scenes/bunny_env.xml')
out = {}
for label, tex, vert, sppe, sppse in (("renderC only", False, False, 0, 0), ("texture grads", True, False, 0, 0), ("texture + vertex grads, interior", True, True, 0, 0),
                                      ("texture + vertex grads, all terms", True, True, SPP, SPP)):
    ctx = capi.Context(0)
    ctx.load_description(desc, dict(width=W, height=H, spp=SPP, sppe=sppe, sppse=sppse))
    if tex:
        for name in ("alpha_u", "alpha_v", "eta", "k"):
            ctx.grad_require(capi.PARAM_BSDF_TEXTURE, 0, name)
        ctx.grad_require(capi.PARAM_ENVMAP_SCALE, 0)
    if vert:
        ctx.grad_require(capi.PARAM_MESH_VERTICES, 0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    t0 = time.time(); ctx.configure(); torch.cuda.synchronize(); t_cfg = time.time() - t0
    integ = capi.make_integrator("direct", bsdf_samples=1, light_samples=1) if depth == 1 else capi.make_integrator("path", max_depth=depth)
    img = torch.empty((W * H, 3), device="cuda"); dLdI = torch.ones_like(img)
    grad = torch.zeros(max(1, ctx.grad_size()), device="cuda")
    res = []
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        ctx.render_c(integ, out=img); torch.cuda.synchronize(); tc = time.time() - t0
        td = tv = 0.0
        if tex or vert:
            t0 = time.time(); ctx.render_d(integ, out=img); torch.cuda.synchronize(); t1 = time.time()
            grad.zero_(); ctx.render_d_vjp(integ, dLdI, grad=grad); torch.cuda.synchronize(); t2 = time.time()
            td, tv = t1 - t0, t2 - t1
        res.append((tc, td, tv))
    tc, td, tv = min(r[0] for r in res), min(r[1] for r in res), min(r[2] for r in res)
    mps = W * H * SPP / (td + tv) / 1e6 if td + tv > 0 else W * H * SPP / tc / 1e6
    print("%-36s configure %.3fs renderC %.3fs renderD %.3fs vjp %.3fs -> %.1f Mpath-samples/s, |grad| %.4g" % (label, t_cfg, tc, td, tv, mps, float(grad.norm())), flush=True)
    out[label] = dict(configure_s=t_cfg, renderC_s=tc, renderD_s=td, vjp_s=tv, Mpath_samples_per_s=mps)
    ctx.close()
print(json.dumps(out))
