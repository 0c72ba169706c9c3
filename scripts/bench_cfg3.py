# BASELINE.json configs[2]: renderD + VJP with vertex-position gradients of the bunny (interior + primary + secondary
# boundary terms) on cbox_bunny at 512x512 / 128 spp (sppe = sppse = 128). Prints timings per stage.
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psdr_cuda_b200 import capi, scene_io
W = H = 512; SPP = 128
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 5
desc = scene_io.load_scene_description('tests/data/scenes/cbox_bunny.xml')
out = {}
for label, sppe, sppse, guide in (("interior only", 0, 0, None), ("+primary edges", SPP, 0, None), ("+secondary edges", 0, SPP, None),
                                  ("all terms", SPP, SPP, None), ("all terms, guided", SPP, SPP, ([40000, 5, 5, 2], 16))):
    ctx = capi.Context(0)
    ctx.load_description(desc, dict(width=W, height=H, spp=SPP, sppe=sppe, sppse=sppse))
    ctx.grad_require(capi.PARAM_MESH_VERTICES, 1)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    t0 = time.time(); ctx.configure(); torch.cuda.synchronize(); t_cfg = time.time() - t0
    integ = capi.make_integrator("path", max_depth=depth, use_guiding=guide is not None)
    t_guide = 0.0
    if guide is not None:
        t0 = time.time(); ctx.preprocess_secondary_edges(0, guide[0], guide[1]); torch.cuda.synchronize(); t_guide = time.time() - t0
    img = torch.empty((W * H, 3), device="cuda"); dLdI = torch.ones_like(img)
    grad = torch.zeros(ctx.grad_size(), device="cuda")
    res = []
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        ctx.render_d(integ, out=img); torch.cuda.synchronize(); t1 = time.time()
        grad.zero_(); ctx.render_d_vjp(integ, dLdI, grad=grad); torch.cuda.synchronize(); t2 = time.time()
        res.append((t1 - t0, t2 - t1))
    td, tv = min(r[0] for r in res), min(r[1] for r in res)
    lanes = W * H * (SPP + sppe + sppse)
    print("%-20s configure %.3fs guide %.2fs renderD %.3fs vjp %.3fs  -> %.1f Mpath-samples/s (interior lanes / (renderD+vjp)), |grad| %.4g" %
          (label, t_cfg, t_guide, td, tv, W * H * SPP / (td + tv) / 1e6, float(grad.norm())), flush=True)
    out[label] = dict(configure_s=t_cfg, guide_s=t_guide, renderD_s=td, vjp_s=tv, Mpath_samples_per_s=W * H * SPP / (td + tv) / 1e6)
    ctx.close()
print(json.dumps(out))
