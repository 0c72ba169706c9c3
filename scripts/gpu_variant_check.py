# renders the same frame with each traversal variant and checks the images agree (hits are exact, film atomics reorder sums)
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psdr_cuda_b200 import capi, scene_io
desc = scene_io.load_scene_description('tests/data/scenes/cbox_bunny.xml')
ctx = capi.Context(0)
ctx.load_description(desc, dict(width=96, height=96, spp=16, sppe=0, sppse=0))
ctx.configure()
integ = capi.make_integrator("path", max_depth=4)
ref = None
for name, kv in (("ld128", dict(trace_ld256=0, trace_sstack=0)), ("ld256", dict(trace_ld256=1, trace_sstack=0)), ("sstack12", dict(trace_sstack=12)),
                 ("sstack16", dict(trace_sstack=16)), ("sstack24", dict(trace_sstack=24))):
    for k, v in kv.items():
        ctx.debug_set(k, v)
    ctx.reset_sampler() if hasattr(ctx, "reset_sampler") else None
    c2 = capi.Context(0); c2.load_description(desc, dict(width=96, height=96, spp=16, sppe=0, sppse=0)); c2.configure()
    img = c2.render_c(integ).clone()
    if ref is None:
        ref = img
    print(name, "max abs diff vs ld128: %.3g" % float((img - ref).abs().max()), "mean %.6f" % float(img.mean()), flush=True)
    assert float((img - ref).abs().max()) < 1e-5
    c2.close()
print("VARIANTS AGREE")
