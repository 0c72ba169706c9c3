#!/usr/bin/env python
"""bench.py — Mpath-samples/s of renderC + renderD(+VJP) (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2|cfg3|cfg4|cfg5]     this repo's sm_100a path
  python bench.py --impl reference [...]      the CPU oracle port timed on the host cores (psdr-cuda itself cannot be built here: SURVEY F4)

Default workload = BASELINE.json configs[1] ("cfg2"): one step = renderC + renderD + its VJP to the diffuse-albedo gradients of
cbox_bunny.xml at 512x512 / 256 spp with the PathIntegrator (max_depth 5); metric = 2*W*H*spp / (t_renderC + t_renderD+vjp).
  cfg3  configs[2]: the same scene at 128 spp, gradients of the bunny's vertex positions, primary + secondary boundary terms (sppe = sppse = 128)
  cfg4  configs[3]: cfg3 over N GPUs (pixel tiles for the interior term, lane ranges for the boundary terms, one all-reduce of the gradient vector)
  cfg5  configs[4]: bunny_env.xml (rough conductor + environment map) at 1024x1024 / 512 spp, texture + vertex gradients, all terms
N > 1: one process per GPU (torchrun). The interior term is split into per-GPU pixel tiles (--shard pixels, default; every rank renders
all samples of its image rows) or by samples (--shard samples); the boundary terms by lane range. The exchange is enqueued by the library
on its own stream with NCCL (csrc/pb_dist.cu): the film of renderC is gathered, and ONE all-reduce of the flat gradient vector ends
renderD + VJP (with pixel tiles the loss is evaluated per tile, so renderD's film needs no exchange). The job is fixed: strong scaling.

The line also carries `verify`: after the timed region the benchmarked sequence is replayed once from freshly seeded samplers and compared
with golden vectors of the CPU oracle at the full benchmark size (tests/golden/bench_cfg2_golden.npz; cfg2) and, on several GPUs, with an
unsharded replay on rank 0 at reduced sample count — the numbers being timed are the right image and the right gradient.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENES = os.path.join(ROOT, "tests", "data", "scenes")
METRIC = "Mpath-samples/s renderC+renderD"
UNIT = "Mpath-samples/s"
BYTES_PER_RAY = 48.0   # traversal: RayRec 32 B read + HitRec 16 B written

CONFIGS = {
    "cfg2": dict(scene="cbox_bunny.xml", w=512, h=512, spp=256, sppe=0, sppse=0, integ=("path", dict(max_depth=5)), grads="albedo",
                 workload="cbox_bunny.xml 512x512/256spp PathIntegrator(max_depth=5) renderC+renderD+VJP, diffuse-albedo gradients"),
    "cfg3": dict(scene="cbox_bunny.xml", w=512, h=512, spp=128, sppe=128, sppse=128, integ=("path", dict(max_depth=5)), grads="bunny_vertices",
                 workload="cbox_bunny.xml 512x512/128spp (sppe=sppse=128) PathIntegrator(max_depth=5) renderC+renderD+VJP, bunny vertex-position gradients, interior + primary + secondary boundary terms"),
    "cfg5": dict(scene="bunny_env.xml", w=1024, h=1024, spp=512, sppe=512, sppse=512, integ=("path", dict(max_depth=3)), grads="rc_textures+vertices",
                 workload="bunny_env.xml 1024x1024/512spp (sppe=sppse=512) PathIntegrator(max_depth=3) renderC+renderD+VJP, rough-conductor texture (alpha_u, alpha_v, eta, k) + envmap scale + bunny vertex gradients, all terms"),
}
CONFIGS["cfg4"] = dict(CONFIGS["cfg3"], workload=CONFIGS["cfg3"]["workload"] + ", sharded over the GPUs")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU arms ---------------------------------------------------------------------------------------------------------------
def host_threads():
    """all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers: that is not a property of the box)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _oracle(threads):
    os.environ["OMP_NUM_THREADS"] = str(threads)     # before the OpenMP runtime of liborc.so starts
    from oracle import orc
    orc.lib().orc_set_num_threads(int(threads))
    return orc


def oracle_sample(cfg, steps, warmup, w=256, h=256, spp=16):
    """Time the CPU oracle (port of the reference algorithm) on a bounded sample of the same workload, on all host threads."""
    import numpy as np
    threads = host_threads()
    orc = _oracle(threads)
    desc = orc.load_scene_description(os.path.join(SCENES, cfg["scene"]))
    sc = orc.Scene(desc, dict(width=w, height=h, spp=spp, sppe=0, sppse=0))
    if cfg["grads"] == "albedo":
        sc.set_bsdf_tangent(0, "reflectance", np.ones((1, 1, 3), np.float32))   # one forward-mode tangent (white albedo)
    else:
        m = 1 if cfg["scene"].startswith("cbox") else 0
        sc.set_mesh_vertex_tangent(m, np.ones_like(desc["meshes"][m]["verts"]))
    sc.configure()
    kind, kw = cfg["integ"]
    integ = orc.PathIntegrator(kw["max_depth"])
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        integ.renderC(sc)
        integ.renderD(sc)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    cores = int(orc.lib().orc_num_threads())
    return dict(value=2.0 * w * h * spp / t / 1e6, t=t, cores=cores,
                sample="%s %dx%d/%dspp PathIntegrator(max_depth=%d): renderC + renderD with one forward-mode tangent (interior term), OpenMP over pixels on %d threads, mean of %d"
                       % (cfg["scene"], w, h, spp, kw["max_depth"], cores, len(times)))


def reference_source_sample(w=128, h=128, spp=16):
    """psdr-cuda's OWN code (oracle/_ref/libref_render.so: its src/**/*.cpp compiled for the CPU against stand-ins for Enoki and OptiX,
    DESIGN.md §2) on the nearest workload it has — the snapshot has no PathIntegrator: DirectIntegrator(1,1) renderC + renderD
    with one forward-mode albedo tangent. Informational: the stand-in is a plain host-array implementation, not a tuned CPU renderer."""
    try:
        from oracle import refrun
        import numpy as np
        if not os.path.exists(refrun.LIB_PATH):
            return None
        refrun.set_matvec_plain(True)
        sc = refrun.Scene(os.path.join(SCENES, "cbox_bunny.xml"), os.path.join(ROOT, "tests"), w, h, spp, 0, 0)
        sc.set_bsdf_tangent(0, "reflectance", np.ones((1, 3), np.float32))
        sc.configure()
        integ = refrun.DirectIntegrator(1, 1)
        t0 = time.perf_counter()
        integ.renderC(sc)
        integ.renderD(sc)
        t = time.perf_counter() - t0
        return {"value": 2.0 * w * h * spp / t / 1e6, "unit": UNIT, "kind": "reference", "seconds": t,
                "sample": "cbox_bunny %dx%d/%dspp DirectIntegrator(1,1): renderC + renderD with one forward-mode tangent, the reference's own source on a CPU stand-in for Enoki/OptiX (single thread + OpenMP ray cast)" % (w, h, spp)}
    except Exception as e:   # the checker library is optional for the bench
        return {"unavailable": str(e)[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    r = oracle_sample(cfg, max(1, args.steps), max(0, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["t"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"] + " (timed on a bounded sample, normalised per path-sample)"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "psdr-cuda has no CPU path and its OptiX/Enoki build is unavailable offline (SURVEY F4): this arm is the CPU oracle port (the only one with a PathIntegrator) on every host thread (%d); reference_source times the reference's own code on its DirectIntegrator" % r["cores"]}
    rs = reference_source_sample()
    if rs is not None:
        line["reference_source"] = rs
    print(json.dumps(line), flush=True)


# ---- this repo's arm ------------------------------------------------------------------------------------------------------------
def require_grads(cfg, desc, ctx, capi):
    if cfg["grads"] == "albedo":
        for b in range(len(desc["bsdfs"])):
            ctx.grad_require(capi.PARAM_BSDF_TEXTURE, b, "reflectance")
    elif cfg["grads"] == "bunny_vertices":
        ctx.grad_require(capi.PARAM_MESH_VERTICES, 1)
    else:
        for name in ("alpha_u", "alpha_v", "eta", "k"):
            ctx.grad_require(capi.PARAM_BSDF_TEXTURE, 0, name)
        ctx.grad_require(capi.PARAM_ENVMAP_SCALE, 0)
        ctx.grad_require(capi.PARAM_MESH_VERTICES, 0)


def module_leaves(cfg, scene):
    """the same parameters as torch leaves of the `psdr_cuda` module (the call a user of the reference makes)"""
    if cfg["grads"] == "albedo":
        return [scene.parameter("BSDF[%d]" % b, "reflectance") for b in range(4)]
    if cfg["grads"] == "bunny_vertices":
        return [scene.parameter("Mesh[1]", "vertex_positions")]
    return [scene.parameter("BSDF[0]", n) for n in ("alpha_u", "alpha_v", "eta", "k")] + [scene.parameter("Emitter[0]", "scale"), scene.parameter("Mesh[0]", "vertex_positions")]


def box64(img, w, h):
    return img.reshape(64, h // 64, 64, w // 64, 3).double().mean(dim=(1, 3)).float()


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from psdr_cuda_b200 import capi, scene_io
    from psdr_cuda_b200 import dist as pdist

    cfg = CONFIGS[args.config]
    W, H, SPP = cfg["w"], cfg["h"], cfg["spp"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    by_pixels = (args.shard == "pixels")

    scene_file = os.path.join(SCENES, cfg["scene"])
    desc = scene_io.load_scene_description(scene_file)
    integ = capi.make_integrator(cfg["integ"][0], **cfg["integ"][1])
    npix = W * H

    def make_ctx(spp, sppe, sppse, sharded=True):
        c = capi.Context(local_rank)
        c.load_description(desc, dict(width=W, height=H, spp=spp, sppe=sppe, sppse=sppse))
        require_grads(cfg, desc, c, capi)
        c.set_stream(torch.cuda.current_stream().cuda_stream)
        if world > 1 and sharded:
            pdist.init_context(c, mode=args.shard, tile_rows=args.tile_rows)     # NCCL communicator owned by the library
        if args.batch:
            c.set_batch(args.batch)
        for kv in args.debug:
            k, v = kv.split("=")
            c.debug_set(k, int(v))
        c.configure()
        return c

    ctx = make_ctx(SPP, cfg["sppe"], cfg["sppse"])
    img_c = torch.empty((npix, 3), dtype=torch.float32, device=dev)
    img_d = torch.empty((npix, 3), dtype=torch.float32, device=dev)
    dLdI = torch.ones((npix, 3), dtype=torch.float32, device=dev)
    grad = torch.zeros(ctx.grad_size(), dtype=torch.float32, device=dev)
    stats = {"trace_ms": 0.0, "rays": 0, "active_rays": 0, "trace_launches": 0, "primary_ms": 0.0}

    def step(collect=False):
        """kernel-level step: everything resident in HBM"""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        ctx.render_c(integ, out=img_c)
        ctx.allreduce_image(img_c)                       # no-op on one GPU
        if collect:
            s = ctx.stats()
            for k in ("trace_ms", "rays", "active_rays", "trace_launches", "primary_ms"):
                stats[k] += s[k]
        ev[1].record()
        ctx.render_d(integ, out=img_d)
        if not by_pixels:
            ctx.allreduce_image(img_d)                   # sample shards hold partial sums of every pixel; pixel tiles keep their own rows
        grad.zero_()
        ctx.render_d_vjp(integ, dLdI, grad=grad)
        ctx.allreduce_grads(grad)                        # the one exchange of renderD: enqueued behind the last adjoint kernel
        ev[2].record()
        return ev

    def timed(fn, steps, warmup, **kw):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        evs = [fn(**kw) for _ in range(steps)]
        t1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = t0.elapsed_time(t1)
        tc = sum(e[0].elapsed_time(e[1]) for e in evs); td = sum(e[1].elapsed_time(e[2]) for e in evs)
        if world > 1:
            t = torch.tensor([ms, tc, td], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, tc, td = t.tolist()
        return ms, tc, td

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0, c0 = ctx.stats()["launches"], ctx.stats()["collectives"]
    ms, tc, td = timed(step, args.steps, args.warmup, collect=True)
    launches, collectives = ctx.stats()["launches"] - l0, ctx.stats()["collectives"] - c0
    launches = launches * args.steps // (args.steps + args.warmup); collectives = collectives * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop() if rank == 0 else None

    # ---- verification of what was timed (untimed): fresh samplers, same sequence, against the CPU oracle's full-size golden vectors
    verify = verify_against_golden(args, cfg, ctx, integ, img_c, img_d, grad, by_pixels, torch, np) if not args.no_verify else None
    ctx.close()
    del ctx
    if world > 1 and not args.no_verify:
        v2 = verify_sharding(args, cfg, make_ctx, integ, world, rank, by_pixels, torch, dist, dev)
        if verify is None:
            verify = {}
        verify["sharded_vs_unsharded"] = v2
    torch.cuda.empty_cache()

    # ---- end to end through the reference-facing plugin: `import psdr_cuda` (pybind11 host module), host buffers in and out
    import psdr_cuda_b200.compat  # noqa: F401
    import psdr_cuda
    scene = psdr_cuda.Scene(local_rank)
    scene.load_file(scene_file, False)
    scene.opts.width, scene.opts.height, scene.opts.spp, scene.opts.sppe, scene.opts.sppse, scene.opts.log_level = W, H, SPP, cfg["sppe"], cfg["sppse"], 0
    leaves = module_leaves(cfg, scene)
    if world > 1:
        scene.init_distributed(mode=args.shard, tile_rows=args.tile_rows)
    scene.configure()
    kind, kw = cfg["integ"]
    m_integ = psdr_cuda.PathIntegrator(kw["max_depth"]) if kind == "path" else psdr_cuda.DirectIntegrator(**kw)
    h_img_c = torch.empty((npix, 3), dtype=torch.float32).pin_memory()
    h_img_d = torch.empty((npix, 3), dtype=torch.float32).pin_memory()
    h_dLdI = torch.ones((npix, 3), dtype=torch.float32).pin_memory()
    h_grads = [torch.empty(p.shape, dtype=torch.float32).pin_memory() for p in leaves]
    d_dLdI = torch.empty((npix, 3), dtype=torch.float32, device=dev)

    def step_e2e():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        ic = m_integ.renderC(scene, 0)                   # Integrator.renderC of the module (film complete on every rank)
        h_img_c.copy_(ic, non_blocking=True)
        ev[1].record()
        idd = m_integ.renderD(scene, 0)                  # torch.autograd node
        h_img_d.copy_(idd.detach(), non_blocking=True)
        d_dLdI.copy_(h_dLdI, non_blocking=True)          # this step's dL/dI arrives from the host
        for p in leaves:
            p.grad = None
        (idd * d_dLdI).sum().backward()                  # pb_render_d_vjp (+ the gradient all-reduce on several GPUs)
        for hg, p in zip(h_grads, leaves):
            hg.copy_(p.grad, non_blocking=True)
        torch.cuda.current_stream().synchronize()        # the step's results are on the host
        ev[2].record()
        return ev

    ms_e, _, _ = timed(step_e2e, args.steps, args.warmup)
    e2e_grad_check = None
    if cfg["grads"] == "albedo" and verify is not None and "grad_all_ones" in verify:
        e2e_grad_check = float(sum(float(hg.double().sum()) for hg in h_grads))   # sum of the 12 albedo gradients for dL/dI = 1, through the module

    lanes_per_rank = W * H * SPP // max(1, world)
    total_samples = 2.0 * W * H * SPP * args.steps
    value = total_samples / (ms * 1e-3) / 1e6
    e2e_value = total_samples / (ms_e * 1e-3) / 1e6
    peak, peak_src = measured_peak()
    # roofline of the dominant kernel (the traversal) from the live CUDA-event durations of the renderC calls of the timed region
    avg_launch_s = stats["trace_ms"] * 1e-3 / max(1, stats["trace_launches"])
    rays_per_launch = stats["active_rays"] / max(1, stats["trace_launches"])   # rays actually traced (unlit / dead lanes are compacted away)
    lanes_per_launch = stats["rays"] / max(1, stats["trace_launches"])
    achieved = BYTES_PER_RAY * rays_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k_trace_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = oracle_sample(cfg, 2, 1)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        par = ("pixel tiles x%d (%s), film gathered for renderC, one all-reduce of the gradient vector per renderD+VJP" % (world, "contiguous row blocks" if args.tile_rows == 0 else "%d-row tiles dealt round-robin" % args.tile_rows)) if by_pixels \
            else "sample-sharded x%d, all-reduce of both films and of the gradient vector" % world
        if world == 1:
            par = "single GPU"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["workload"], "name": args.config,
                           "scene": "tests/data/scenes/%s (the reference ships no plain cbox.xml, SURVEY F2)" % cfg["scene"],
                           "l2": "wavefront buffers of one batch exceed L2 (%.0f MB per batch) and are rewritten every batch" % (ctx_batch_mb(args)),
                           "parallelism": par, "collectives_per_step": int(collectives // max(1, args.steps)),
                           "ms_renderC": tc / args.steps, "ms_renderD_vjp": td / args.steps,
                           "Mpath_samples_per_s_renderC": W * H * SPP * args.steps / (tc * 1e-3) / 1e6,
                           "Mpath_samples_per_s_renderD_vjp": W * H * SPP * args.steps / (td * 1e-3) / 1e6,
                           "boundary_lanes_per_step": W * H * (cfg["sppe"] + cfg["sppse"])},
                "roofline": {"bound": "hbm", "kernel": "k_trace_stream", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_ray": BYTES_PER_RAY,
                             "rays_per_launch": rays_per_launch, "ray_slots_per_launch": lanes_per_launch, "avg_launch_ms": avg_launch_s * 1e3,
                             "Grays_per_s": rays_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0,
                             "note": "BVH traversal without RT cores is bound by instruction issue, the ALU pipe and L1 (ncu: profiles/), not by HBM bytes",
                             "overlapped": bool(lanes_per_rank <= (20 << 20)),   # short renders run two half-batches on two streams: the per-launch durations then include the other stream's kernels
                             },
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_dLdI.numel() * 4),
                        "d2h_bytes_per_step": int((h_img_c.numel() + h_img_d.numel() + sum(g.numel() for g in h_grads)) * 4),
                        "through": "import psdr_cuda (pybind11 host module): Integrator.renderC / renderD + torch.autograd backward, pinned host buffers",
                        "grad_sum_for_unit_dLdI": e2e_grad_check},
                "verify": verify,
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        scene.dist_finalize()
        dist.destroy_process_group()


def verify_against_golden(args, cfg, ctx, integ, img_c, img_d, grad, by_pixels, torch, np):
    """replay renderC / renderD / VJP from freshly seeded samplers and compare with the CPU oracle's run of the same (full-size) job"""
    gp = os.path.join(ROOT, "tests", "golden", "bench_%s_golden.npz" % ("cfg2" if args.config == "cfg2" else args.config))
    if not os.path.exists(gp) or args.config != "cfg2":
        return {"golden": None, "note": "no full-size oracle golden for this configuration (the oracle needs hours); parity of its terms is covered at test size by tests/test_gpu_parity.py"}
    G = np.load(gp)
    W, H = cfg["w"], cfg["h"]
    dev = img_c.device
    out = {"golden": "tests/golden/bench_cfg2_golden.npz (CPU oracle, 512x512/256spp PathIntegrator(5), tests/golden/make_bench_golden.py)"}
    ctx.configure(reseed=True)
    ctx.render_c(integ, out=img_c); ctx.allreduce_image(img_c)
    c64 = box64(img_c, W, H).cpu().numpy()
    out["renderC_mean"] = float(img_c.double().mean()); out["renderC_mean_golden"] = float(G["meanC"])
    out["renderC_box64_max_rel_err"] = float(np.abs(c64 - G["C64"]).max() / np.abs(G["C64"]).max())
    ctx.configure(reseed=True)
    ctx.render_d(integ, out=img_d); ctx.allreduce_image(img_d)
    d64 = box64(img_d, W, H).cpu().numpy()
    out["renderD_mean"] = float(img_d.double().mean()); out["renderD_mean_golden"] = float(G["meanD"])
    out["renderD_box64_max_rel_err"] = float(np.abs(d64 - G["D64"]).max() / np.abs(G["D64"]).max())
    y, x = np.mgrid[0:H, 0:W]
    ramp = torch.from_numpy(np.stack([x / W, y / H, np.ones_like(x, dtype=np.float64)], -1).reshape(-1, 3).astype(np.float32)).to(dev)
    errs = []
    for p, dl in enumerate((torch.ones_like(img_d), ramp)):
        grad.zero_()
        ctx.render_d_vjp(integ, dl, grad=grad); ctx.allreduce_grads(grad)
        g = grad.double().cpu().numpy()
        if p == 0:
            out["grad_all_ones"] = float(g.sum())
        for k in range(G["tangents"].shape[0]):
            got = float((g * G["tangents"][k].astype(np.float64)).sum()); want = float(G["proj"][k, p])
            errs.append(abs(got - want) / abs(want))
    out["grad_projection_rel_err"] = errs      # <dL/dI_p, dI/dtheta . t_k> for two directions t_k over the 12 albedo entries and two dL/dI
    out["ok"] = bool(out["renderC_box64_max_rel_err"] <= 1e-3 and out["renderD_box64_max_rel_err"] <= 1e-3 and max(errs) <= 1e-3
                     and abs(out["renderC_mean"] - out["renderC_mean_golden"]) <= 1e-4 * abs(out["renderC_mean_golden"]))
    return out


def verify_sharding(args, cfg, make_ctx, integ, world, rank, by_pixels, torch, dist, dev):
    """N ranks vs one: the sharded job (reduced sample counts) against an unsharded replay on rank 0"""
    W, H = cfg["w"], cfg["h"]
    spp = max(world, cfg["spp"] // 16)
    sppe, sppse = (max(1, cfg["sppe"] // 16) if cfg["sppe"] else 0), (max(1, cfg["sppse"] // 16) if cfg["sppse"] else 0)
    c = make_ctx(spp, sppe, sppse)
    img = torch.empty((W * H, 3), dtype=torch.float32, device=dev); imgd = torch.empty_like(img)
    g = torch.zeros(c.grad_size(), dtype=torch.float32, device=dev)
    c.render_c(integ, out=img); c.allreduce_image(img)
    c.render_d(integ, out=imgd); c.allreduce_image(imgd)
    c.render_d_vjp(integ, torch.ones_like(img), grad=g); c.allreduce_grads(g)
    torch.cuda.synchronize()
    c.close()
    out = None
    if rank == 0:
        c1 = make_ctx(spp, sppe, sppse, sharded=False)
        img1 = torch.empty_like(img); imgd1 = torch.empty_like(img)
        g1 = torch.zeros_like(g)
        c1.render_c(integ, out=img1); c1.render_d(integ, out=imgd1); c1.render_d_vjp(integ, torch.ones_like(img), grad=g1)
        torch.cuda.synchronize()
        c1.close()
        gn = float(g1.double().norm())
        out = {"spp": spp, "sppe": sppe, "sppse": sppse,
               "renderC_max_abs_diff": float((img - img1).abs().max()), "renderD_max_abs_diff": float((imgd - imgd1).abs().max()),
               "grad_rel_l2": float((g.double() - g1.double()).norm() / max(gn, 1e-30)),
               "bit_identical_films": bool(torch.equal(img, img1) and torch.equal(imgd, imgd1))}
        out["ok"] = bool(out["renderC_max_abs_diff"] <= 1e-4 and out["renderD_max_abs_diff"] <= 1e-4 and out["grad_rel_l2"] <= 1e-3)
    dist.barrier()
    return out


def ctx_batch_mb(args):
    lanes = args.batch if args.batch else (1 << 25)
    return lanes * (16 + 2 * 2 * (32 + 16) + 2 * 32) / 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default cfg2 = configs[1], the one the metric is quoted on)")
    ap.add_argument("--shard", default="pixels", choices=["pixels", "samples"], help="N > 1: how the interior term is split over the GPUs")
    ap.add_argument("--tile-rows", type=int, default=4, help="pixel sharding: rows per tile, dealt round-robin (0 = one contiguous block per GPU: 3.7 %% load imbalance on two GPUs against 1.4 %% with 8-row tiles; 4 rows: +1.7 %% on eight GPUs)")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--debug", action="append", default=[], help="debug: key=value passed to pb_debug_set")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
