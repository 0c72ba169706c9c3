import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psdr_cuda_b200 import capi, scene_io
variants = [int(v) for v in sys.argv[1].split(",")]
desc = scene_io.load_scene_description('tests/data/scenes/cbox_bunny.xml')
ctx = capi.Context(0)
ctx.load_description(desc, dict(width=512, height=16, spp=256, sppe=0, sppse=0))
ctx.configure()
B = 1 << 20
ctx.set_batch(B)
ctx.debug_set("trace_variant", 0)
ctx.render_c(capi.make_integrator("path", max_depth=2))
ptr, nbytes = ctx.debug_ray_buffer(1)
t = torch.empty((2 * B, 8), dtype=torch.float32, device="cuda")
ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(2 * B * 32), 3)
torch.cuda.synchronize()
for v in variants:
    ctx.debug_set("trace_variant", v)
    for _ in range(2):
        ctx.trace(t)
