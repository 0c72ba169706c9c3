# Scene::configure cost inside an optimisation loop (vertex edits every iteration), cbox_bunny at the cfg3 sample counts:
# host SAH rebuild vs device refit, with and without the edge tables of the boundary terms.
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psdr_cuda_b200 import capi, scene_io
desc = scene_io.load_scene_description('tests/data/scenes/cbox_bunny.xml')
rng = np.random.default_rng(0)
for label, sppe, sppse in (("interior only", 0, 0), ("with primary + secondary edge tables", 128, 128)):
    for refit in (0, 16):
        ctx = capi.Context(0)
        ctx.load_description(desc, dict(width=512, height=512, spp=128, sppe=sppe, sppse=sppse))
        ctx.grad_require(capi.PARAM_MESH_VERTICES, 1)
        ctx.set_bvh_refit(refit)
        ctx.configure()
        verts = desc["meshes"][1]["verts"].copy()
        ts = []
        for it in range(5):
            verts = verts + rng.normal(scale=0.01, size=verts.shape).astype(np.float32)
            ctx.set_mesh_vertices(1, verts)
            torch.cuda.synchronize(); t0 = time.time()
            ctx.configure()
            torch.cuda.synchronize(); ts.append(time.time() - t0)
        print("%-40s %s: configure %.1f ms (min of 5), bvh %s" % (label, "refit  " if refit else "rebuild", 1e3 * min(ts), ctx.bvh_stats()), flush=True)
        ctx.close()

# cfg5's scene: rough conductor + environment map (1024 x 512 radiance map -> 2 M-cell distribution), envmap scale edited every iteration
desc5 = scene_io.load_scene_description('tests/data/scenes/bunny_env.xml')
for label, edit_env in (("bunny_env, vertex edit, all tables", False), ("bunny_env, vertex + envmap radiance edit", True)):
    ctx = capi.Context(0)
    ctx.load_description(desc5, dict(width=1024, height=1024, spp=64, sppe=64, sppse=64))
    ctx.grad_require(capi.PARAM_MESH_VERTICES, 0)
    torch.cuda.synchronize(); t0 = time.time()
    ctx.configure()
    torch.cuda.synchronize(); t_first = time.time() - t0
    verts = desc5["meshes"][0]["verts"].copy()
    rad = desc5["envmap"]["radiance"].copy()
    ts = []
    for it in range(5):
        verts = verts + rng.normal(scale=0.001, size=verts.shape).astype(np.float32)
        ctx.set_mesh_vertices(0, verts)
        if edit_env:
            rad = rad * np.float32(1.01)
            ctx.set_envmap_radiance(rad, desc5["envmap"]["scale"])
        torch.cuda.synchronize(); t0 = time.time()
        ctx.configure()
        torch.cuda.synchronize(); ts.append(time.time() - t0)
    print("%-45s first configure %.1f ms, then %.1f ms (min of 5), bvh %s" % (label, 1e3 * t_first, 1e3 * min(ts), ctx.bvh_stats()), flush=True)
    ctx.close()
