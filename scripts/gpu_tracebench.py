import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psdr_cuda_b200 import capi, scene_io
variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2,3").split(",")]
desc = scene_io.load_scene_description('tests/data/scenes/cbox_bunny.xml')
ctx = capi.Context(0)
ctx.load_description(desc, dict(width=512, height=16, spp=256, sppe=0, sppse=0))
for b in range(4): ctx.grad_require(capi.PARAM_BSDF_TEXTURE, b, "reflectance")
ctx.configure()
B = 1 << 20
ctx.set_batch(B)
bufs = []
for k in range(5):
    ctx.configure(reseed=True)
    ctx.render_c(capi.make_integrator("path", max_depth=k + 1))
    ptr, nbytes = ctx.debug_ray_buffer(k)
    n = 2 * B
    t = torch.empty((n, 8), dtype=torch.float32, device="cuda")
    ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(n * 32), 3)
    bufs.append(t)
    print("event", k, "active fraction bsdf %.3f light %.3f" % ((t[:B, 3] > 0).float().mean().item(), (t[B:, 3] > 0).float().mean().item()))
ref = {}
runs = []
for v in variants:
    if v == 4:
        runs += [(4, 8)]
    else:
        runs.append((v, 8))
for v, bps in runs:
    ctx.debug_set("trace_variant", v)
    ctx.debug_set("trace_blocks_per_sm", bps)
    line = "variant %d bps %d:" % (v, bps)
    for k, t in enumerate(bufs):
        for part, sl in (("b", slice(0, B)), ("l", slice(B, 2 * B)), ("both", slice(0, 2 * B))):
            r = t[sl].contiguous()
            hits, tt = ctx.trace(r)
            key = (k, part)
            if key not in ref: ref[key] = (hits.clone(), tt.clone())
            elif v < 5: assert torch.equal(hits, ref[key][0]) and torch.equal(tt.view(torch.int32), ref[key][1].view(torch.int32)), (v, key)
            else:   # occlusion queries may stop at a different (closer-than-t_occ) hit: compare where no early-out applies
                same = (hits == ref[key][0]).all(dim=1)
                occl = (tt <= r[:, 7]) & (r[:, 7] > 0)
                assert bool((same | occl).all()), (v, key, int((~(same | occl)).sum()))
            ms = []
            for _ in range(5):
                ctx.trace(r); ms.append(ctx.stats()["trace_ms"])
            line += " e%d%s %.3fms(%.2fG/s)" % (k, part, min(ms), r.shape[0] / min(ms) / 1e6)
    print(line)
print("DONE")
