#!/usr/bin/env python
"""bench.py — Mpath-samples/s of renderC + renderD(+VJP) on the Cornell-box scene (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's sm_100a path through the C ABI
  python bench.py --impl reference [...]                         the CPU oracle port timed on the host cores
                                                                 (psdr-cuda itself cannot be built here: SURVEY F4)

One step = renderC + renderD + its VJP to the diffuse-albedo gradients of cbox_bunny.xml at 512x512 / 256 spp with the
PathIntegrator (max_depth 5). metric = 2*W*H*spp / (t_renderC + t_renderD+vjp) in Mpath-samples/s.
N > 1: one process per GPU (torchrun), each rank renders spp/N samples of every pixel; the image and the flat gradient
vector are all-reduced over NCCL (strong scaling: the job is fixed).
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

SCENE = os.path.join(ROOT, "tests", "data", "scenes", "cbox_bunny.xml")
W, H, SPP, DEPTH = 512, 512, 256, 5
METRIC = "Mpath-samples/s renderC+renderD"
UNIT = "Mpath-samples/s"
BYTES_PER_RAY = 48.0   # k_trace: RayRec 32 B read + HitRec 16 B written


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


def oracle_sample(steps, warmup, w=256, h=256, spp=16):
    """Time the CPU oracle (port of the reference algorithm) on a bounded sample of the same workload."""
    from oracle import orc
    desc = orc.load_scene_description(SCENE)
    sc = orc.Scene(desc, dict(width=w, height=h, spp=spp, sppe=0, sppse=0))
    import numpy as np
    sc.set_bsdf_tangent(0, "reflectance", np.ones((1, 1, 3), np.float32))   # one forward-mode tangent (white albedo)
    sc.configure()
    integ = orc.PathIntegrator(DEPTH)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        integ.renderC(sc)
        integ.renderD(sc)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    cores = orc.lib().orc_num_threads()
    return dict(value=2.0 * w * h * spp / t / 1e6, t=t, cores=int(cores),
                sample="cbox_bunny %dx%d/%dspp PathIntegrator(max_depth=%d): renderC + renderD with one forward-mode tangent, OpenMP over pixels, mean of %d" % (w, h, spp, DEPTH, len(times)))


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
        sc = refrun.Scene(SCENE, os.path.join(ROOT, "tests"), w, h, spp, 0, 0)
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
    r = oracle_sample(max(1, args.steps), max(0, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["t"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cbox_bunny.xml 512x512/256spp PathIntegrator(max_depth=5) renderC+renderD, diffuse-albedo gradients (timed on a bounded sample, normalised per path-sample)"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "psdr-cuda has no CPU path and its OptiX/Enoki build is unavailable offline (SURVEY F4): this arm is the CPU oracle port (the only one with a PathIntegrator); reference_source times the reference's own code on its DirectIntegrator"}
    rs = reference_source_sample()
    if rs is not None:
        line["reference_source"] = rs
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from psdr_cuda_b200 import capi, scene_io

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

    desc = scene_io.load_scene_description(SCENE)
    ctx = capi.Context(local_rank)
    ctx.load_description(desc, dict(width=W, height=H, spp=SPP, sppe=0, sppse=0))
    nb = len(desc["bsdfs"])
    for b in range(nb):
        ctx.grad_require(capi.PARAM_BSDF_TEXTURE, b, "reflectance")
    ctx.set_shard(rank, world)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if args.batch:
        ctx.set_batch(args.batch)
    if args.trace_variant >= 0:
        ctx.debug_set("trace_variant", args.trace_variant)
    for kv in args.debug:
        k, v = kv.split("=")
        ctx.debug_set(k, int(v))
    ctx.configure()
    integ = capi.make_integrator("path", max_depth=DEPTH)
    npix = W * H
    img_c = torch.empty((npix, 3), dtype=torch.float32, device=dev)
    img_d = torch.empty((npix, 3), dtype=torch.float32, device=dev)
    dLdI = torch.ones((npix, 3), dtype=torch.float32, device=dev)
    grad = torch.zeros(ctx.grad_size(), dtype=torch.float32, device=dev)
    h_img_c = torch.empty((npix, 3), dtype=torch.float32).pin_memory()
    h_img_d = torch.empty((npix, 3), dtype=torch.float32).pin_memory()
    h_dLdI = torch.ones((npix, 3), dtype=torch.float32).pin_memory()
    h_grad = torch.empty(ctx.grad_size(), dtype=torch.float32).pin_memory()

    stats = {"trace_ms": 0.0, "rays": 0, "active_rays": 0, "trace_launches": 0, "primary_ms": 0.0, "tc": 0.0, "td": 0.0}

    def step(e2e, collect=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        ctx.render_c(integ, out=img_c)
        if world > 1:
            dist.all_reduce(img_c)
        if e2e:
            h_img_c.copy_(img_c, non_blocking=True)
        if collect:
            s = ctx.stats()
            for k in ("trace_ms", "rays", "active_rays", "trace_launches", "primary_ms"):
                stats[k] += s[k]
        ev[1].record()
        ctx.render_d(integ, out=img_d)
        if world > 1:
            dist.all_reduce(img_d)
        if e2e:
            h_img_d.copy_(img_d, non_blocking=True)
            dLdI.copy_(h_dLdI, non_blocking=True)
        grad.zero_()
        ctx.render_d_vjp(integ, dLdI, grad=grad)
        if world > 1:
            dist.all_reduce(grad)
        if e2e:
            h_grad.copy_(grad, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        ev[2].record()
        return ev

    def timed(e2e, collect):
        for _ in range(args.warmup):
            step(e2e)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = ctx.stats()["launches"]
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        evs = [step(e2e, collect) for _ in range(args.steps)]
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
        return ms, tc, td, ctx.stats()["launches"] - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, tc, td, launches = timed(False, True)
    clocks = sampler.stop() if rank == 0 else None
    ms_e, _, _, _ = timed(True, False)

    total_samples = 2.0 * W * H * SPP * args.steps
    value = total_samples / (ms * 1e-3) / 1e6
    e2e_value = total_samples / (ms_e * 1e-3) / 1e6
    peak, peak_src = measured_peak()
    # roofline of the dominant kernel (k_trace) from the live CUDA-event durations of the renderC calls of the timed region
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
            r = oracle_sample(2, 1)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "cbox_bunny.xml 512x512/256spp PathIntegrator(max_depth=5) renderC+renderD+VJP, diffuse-albedo gradients",
                           "scene": "tests/data/scenes/cbox_bunny.xml (69642 triangles; the reference ships no plain cbox.xml, SURVEY F2)",
                           "l2": "wavefront buffers of one batch exceed L2 (%.0f MB per batch) and are rewritten every batch" % (ctx_batch_mb(args)),
                           "parallelism": "sample-sharded x%d, all-reduce of image and gradient vector" % world,
                           "ms_renderC": tc / args.steps, "ms_renderD_vjp": td / args.steps,
                           "Mpath_samples_per_s_renderC": W * H * SPP * args.steps / (tc * 1e-3) / 1e6,
                           "Mpath_samples_per_s_renderD_vjp": W * H * SPP * args.steps / (td * 1e-3) / 1e6},
                "roofline": {"bound": "hbm", "kernel": "k_trace_perm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_ray": BYTES_PER_RAY,
                             "rays_per_launch": rays_per_launch, "ray_slots_per_launch": lanes_per_launch, "avg_launch_ms": avg_launch_s * 1e3,
                             "Grays_per_s": rays_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0,
                             "note": "traversal is L2-latency/divergence bound (no RT cores on B200); scene tables are L2 resident"},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_dLdI.numel() * 4), "d2h_bytes_per_step": int((h_img_c.numel() + h_img_d.numel() + h_grad.numel()) * 4)},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ctx_batch_mb(args):
    lanes = args.batch if args.batch else (1 << 25)
    return lanes * (16 + 2 * 2 * (32 + 16) + 2 * 32) / 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--trace-variant", type=int, default=-1, help="debug: traversal kernel variant (default: library default)")
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
