import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import numpy as np, torch
from oracle import orc
from psdr_cuda_b200 import capi

desc = orc.load_scene_description('tests/data/scenes/cbox_bunny.xml')
opts = dict(width=128, height=128, spp=16, sppe=0, sppse=0)
osc = orc.Scene(desc, opts); osc.configure()
ctx = capi.Context(0); ctx.load_description(desc, opts); ctx.configure()
ti_o, ti_p = osc.triangle_info(), ctx.triangle_info()
print("tri table: bit-equal =", np.array_equal(ti_o.view(np.uint32), ti_p.view(np.uint32)), "max abs diff", np.abs(ti_o - ti_p).max())
bad = np.argwhere(ti_o.view(np.uint32) != ti_p.view(np.uint32))
print("mismatching entries:", len(bad), bad[:5])
for m in range(len(desc['meshes'])):
    assert np.array_equal(osc.mesh_edges(m), ctx.mesh_edges(m)), m
print("edges equal")
# trace parity on random rays
rng = np.random.default_rng(1)
n = 200000
o = np.zeros((n, 3), np.float32); o[:, 0] = rng.uniform(-90, 90, n); o[:, 1] = rng.uniform(5, 190, n); o[:, 2] = rng.uniform(-90, 190, n)
d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
rays = np.zeros((n, 8), np.float32); rays[:, :3] = o; rays[:, 3] = np.inf; rays[:, 4:7] = d
tri, shape, u, v, t = osc.trace(o, d)
hits, tt = ctx.trace(torch.from_numpy(rays).cuda())
hits = hits.cpu().numpy(); tt = tt.cpu().numpy()
print("trace: tri equal", np.mean(hits[:, 0] == tri), "shape equal", np.mean(hits[:, 1] == shape),
      "uv bit-equal", np.mean(hits[:, 2].view(np.float32).view(np.uint32) == u.view(np.uint32)), np.mean(hits[:, 3].view(np.uint32) == v.view(np.uint32)),
      "t bit-equal", np.mean(tt.view(np.uint32) == t.view(np.uint32)), "hit frac", np.mean(tri >= 0))
# renderC parity cfg1
for kind, kw in (("direct", dict(bsdf_samples=1, light_samples=1)), ("path", dict(max_depth=1)), ("path", dict(max_depth=3)), ("direct", dict(bsdf_samples=2, light_samples=2))):
    osc = orc.Scene(desc, opts); osc.configure()
    ctx.configure(reseed=True)
    if kind == "direct":
        oi = orc.DirectIntegrator(kw['bsdf_samples'], kw['light_samples'])
    else:
        oi = orc.PathIntegrator(kw['max_depth'])
    pi = capi.make_integrator(kind, **kw)
    t0 = time.time(); ref = oi.renderC(osc); t_cpu = time.time() - t0
    torch.cuda.synchronize(); t0 = time.time(); img = ctx.render_c(pi).cpu().numpy(); t_gpu = time.time() - t0
    diff = np.abs(img - ref).mean(axis=1)
    print(kind, kw, "cpu %.2fs gpu %.4fs" % (t_cpu, t_gpu), "mean", ref.mean(), img.mean(), "pixel L1: max %.3e mean %.3e frac>1e-4 %.4f" % (diff.max(), diff.mean(), np.mean(diff > 1e-4)))
    # second call continues the stream
    ref2 = oi.renderC(osc); img2 = ctx.render_c(pi).cpu().numpy()
    d2 = np.abs(img2 - ref2).mean(axis=1)
    print("   second call: mean %.3e frac>1e-4 %.4f ; differs from first: %s" % (d2.mean(), np.mean(d2 > 1e-4), np.abs(img2 - img).mean()))
np.save('gpurun_out/cfg1_gpu.npy', img)
# timing at larger sizes
for (w, h, spp, kind, kw) in ((512, 512, 64, "direct", dict(bsdf_samples=1, light_samples=1)), (512, 512, 256, "direct", dict(bsdf_samples=1, light_samples=1)), (512, 512, 256, "path", dict(max_depth=5))):
    ctx.set_options(w, h, spp, 0, 0); ctx.configure(reseed=True)
    pi = capi.make_integrator(kind, **kw)
    ctx.render_c(pi)
    for batch in (1 << 18, 1 << 20, 1 << 22):
        ctx.set_batch(batch)
        torch.cuda.synchronize(); t0 = time.time(); img = ctx.render_c(pi); torch.cuda.synchronize(); dt = time.time() - t0
        st = ctx.stats()
        print(w, h, spp, kind, kw, "batch", batch, "time %.4fs  Msamples/s %.1f  trace_ms %.2f rays %d Grays/s(trace only) %.2f" % (dt, w * h * spp / dt / 1e6, st['trace_ms'], st['rays'], st['rays'] / st['trace_ms'] / 1e6))
print("DONE")
