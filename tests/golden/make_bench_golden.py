"""Golden vectors of the BENCHMARKED configuration (BASELINE.json configs[1]: cbox_bunny 512x512 / 256 spp, PathIntegrator(5)),
from the CPU oracle at full size. bench.py replays the same sequence on the GPU (untimed) and prints the comparison in its JSON line:

    reseed -> renderC                      -> C64   (image box-filtered to 64x64), meanC
    reseed -> renderD (+ forward tangent)  -> D64, meanD, proj[k, p] = <dLdI_p, dI/d(albedo) . t_k>

t_k: directions over the 12 diffuse-albedo parameters (k = 0: all ones, k = 1: fixed signs); dLdI_p: p = 0 ones, p = 1 the ramp
(x / W, y / H, 1) per channel. Takes about 20 minutes on 8 cores:  python tests/golden/make_bench_golden.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

W, H, SPP, DEPTH = 512, 512, 256, 5
SCENE = os.path.join(ROOT, "tests", "data", "scenes", "cbox_bunny.xml")
TANGENTS = np.array([[1.0] * 12, [1, -1, 1, 1, 1, -1, -1, 1, 1, 1, -1, -1]], np.float32)


def ramp(w=W, h=H):
    y, x = np.mgrid[0:h, 0:w]
    return np.stack([x / w, y / h, np.ones_like(x, dtype=np.float64)], -1).reshape(-1, 3)


def box64(img, w=W, h=H):
    return img.reshape(64, h // 64, 64, w // 64, 3).astype(np.float64).mean(axis=(1, 3)).astype(np.float32)


def main(w=W, h=H, spp=SPP, out=None):
    opts = dict(width=w, height=h, spp=spp, sppe=0, sppse=0)
    desc = orc.load_scene_description(SCENE)
    integ = orc.PathIntegrator(DEPTH)
    t0 = time.time()
    sc = orc.Scene(desc, opts); sc.configure()
    imgC = integ.renderC(sc)
    print("renderC %.0f s" % (time.time() - t0), flush=True)
    res = dict(C64=box64(imgC, w, h), meanC=np.float64(imgC.astype(np.float64).mean()), tangents=TANGENTS, size=np.array([w, h, spp, DEPTH]))
    dL = [np.ones((w * h, 3)), ramp(w, h)]
    proj = np.zeros((len(TANGENTS), 2))
    for k, t in enumerate(TANGENTS):
        sc = orc.Scene(desc, opts)
        for b in range(4):
            sc.set_bsdf_tangent(b, "reflectance", t[3 * b:3 * b + 3].reshape(1, 1, 3))
        sc.configure()
        imgD, dimg = integ.renderD(sc)
        for p in range(2):
            proj[k, p] = float((dL[p] * dimg.astype(np.float64)).sum())
        if k == 0:
            res["D64"] = box64(imgD, w, h); res["meanD"] = np.float64(imgD.astype(np.float64).mean())
        print("renderD tangent %d: %.0f s, proj %s" % (k, time.time() - t0, proj[k]), flush=True)
    res["proj"] = proj
    out = out or os.path.join(ROOT, "tests", "golden", "bench_cfg2_golden.npz")
    np.savez_compressed(out, **res)
    print("wrote", out)


if __name__ == "__main__":
    if len(sys.argv) > 1:   # reduced size for a quick self-check: w h spp out
        main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
    else:
        main()
