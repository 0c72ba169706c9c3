"""GPU test (needs two GPUs on the box; skipped otherwise): two ranks render for real. Each rank owns one process, one context and the
library's own NCCL communicator (csrc/pb_dist.cu); the sharded renderC / renderD / VJP (pixel tiles and sample shards, interior and
boundary terms) must reproduce the single-GPU film and gradient, with ONE collective for renderD + VJP under pixel tiles."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import json, os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from psdr_cuda_b200 import capi, scene_io
from psdr_cuda_b200 import dist as pdist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
desc = scene_io.load_scene_description(os.path.join(%(root)r, "tests", "data", "scenes", "cbox_bunny.xml"))
opts = dict(width=64, height=64, spp=8, sppe=4, sppse=4)
integ = capi.make_integrator("path", max_depth=3)
dLdI = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, size=(64 * 64, 3)).astype(np.float32)).cuda()
def run(mode, tile_rows, sharded):
    c = capi.Context(rank)
    c.load_description(desc, opts)
    for b in range(4): c.grad_require(capi.PARAM_BSDF_TEXTURE, b, "reflectance")
    c.grad_require(capi.PARAM_MESH_VERTICES, 1)
    if sharded: pdist.init_context(c, mode=mode, tile_rows=tile_rows)
    c.configure()
    n0 = c.stats()["collectives"]
    ic = c.render_c(integ); c.allreduce_image(ic)
    n1 = c.stats()["collectives"]
    idd = c.render_d(integ)
    if mode != "pixels": c.allreduce_image(idd)
    g = c.render_d_vjp(integ, dLdI); c.allreduce_grads(g)
    n2 = c.stats()["collectives"]
    if mode == "pixels" and sharded: c.allreduce_image(idd)      # only to compare the film here
    torch.cuda.synchronize()
    out = (ic.cpu().numpy(), idd.cpu().numpy(), g.cpu().numpy().astype(np.float64), n1 - n0, n2 - n1)
    c.close()
    return out
res = {}
ref = run("samples", 0, False) if rank == 0 else None
for name, mode, tile in (("pixels_blocks", "pixels", 0), ("pixels_interleaved", "pixels", 1), ("samples", "samples", 0)):
    ic, idd, g, nc_c, nc_d = run(mode, tile, True)
    if rank == 0:
        res[name] = dict(c_max=float(np.abs(ic - ref[0]).max()), d_max=float(np.abs(idd - ref[1]).max()), c_equal=bool(np.array_equal(ic, ref[0])),
                         g_rel=float(np.linalg.norm(g - ref[2]) / np.linalg.norm(ref[2])), collectives_c=int(nc_c), collectives_d=int(nc_d))
# the module surface: import psdr_cuda, sharded scene, torch.autograd backward
import psdr_cuda_b200.compat, psdr_cuda
def module_run(sharded):
    sc = psdr_cuda.Scene(rank); sc.load_file(os.path.join(%(root)r, "tests", "data", "scenes", "cbox_bunny.xml"), False)
    sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse, sc.opts.log_level = 64, 64, 8, 0, 0, 0
    p = sc.parameter("BSDF[id=white]", "reflectance")
    if sharded: sc.init_distributed()
    sc.configure()
    it = psdr_cuda.PathIntegrator(3)
    img = it.renderD(sc, 0)
    (img * dLdI).sum().backward()
    return img.detach().cpu().numpy(), p.grad.cpu().numpy().astype(np.float64)
mi, mg = module_run(True)
if rank == 0:
    ri, rg = module_run(False)
    res["module"] = dict(img_max=float(np.abs(mi - ri).max()), g_rel=float(np.linalg.norm(mg - rg) / np.linalg.norm(rg)))
dist.barrier()
if rank == 0: print(json.dumps(res))
dist.destroy_process_group()
'''


def test_two_ranks_render_the_single_gpu_result(native_lib, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = 29600 + (os.getpid() % 1000)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    r = json.loads([ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")][-1])
    for name in ("pixels_blocks", "pixels_interleaved", "samples"):
        assert r[name]["c_max"] <= 1e-5 and r[name]["d_max"] <= 1e-5, r
        assert r[name]["g_rel"] <= 1e-3, r                     # gradients: atomics order differs between one and two GPUs
    assert r["pixels_blocks"]["c_equal"] and r["pixels_interleaved"]["c_equal"]       # disjoint pixel tiles: the film is bit-identical
    assert r["pixels_blocks"]["collectives_c"] == 1 and r["pixels_blocks"]["collectives_d"] == 1     # film gather; ONE gradient all-reduce for renderD + VJP
    assert r["module"]["img_max"] <= 1e-5 and r["module"]["g_rel"] <= 1e-3, r
