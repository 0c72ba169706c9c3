"""BASELINE.json's configurations at reduced (but not tiny) sizes through psdr-cuda's own source (oracle/_ref/libref_render.so) and through
the oracle, on the CPU: outlier pixels, projections, wall-clock. Evidence log: profiles/r01f_ref_source_vs_oracle_configs.log

    python scripts/cpu_ref_vs_oracle_configs.py > profiles/r01f_ref_source_vs_oracle_configs.log
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc, refrun  # noqa: E402

T = os.path.join(ROOT, "tests")
refrun.set_matvec_plain(True)


def stats(a, b, rel):
    scale = max(np.abs(b).max(), 1e-12)
    d = np.abs(a - b).max(axis=1)
    nz = int(((np.abs(a).max(axis=1) > 0) | (np.abs(b).max(axis=1) > 0)).sum())
    return "max %.4g, %d of %d non-zero px over %g of max, sum ref %.6g orc %.6g (rel %.2e)" % (
        scale, int((d > rel * scale).sum()), nz, rel, a.astype(np.float64).sum(), b.astype(np.float64).sum(),
        abs(a.astype(np.float64).sum() - b.astype(np.float64).sum()) / max(abs(b.astype(np.float64).sum()), 1e-12))


def run(label, name, opts, leaves, guiding=None, bs=1, ls=1):
    w, h, spp, sppe, sppse = opts
    xml = os.path.join(T, "data", "scenes", name + ".xml")
    rng = np.random.default_rng(17)
    r = refrun.Scene(xml, T, w, h, spp, sppe, sppse)
    d = orc.load_scene_description(xml)
    o = orc.Scene(d, dict(width=w, height=h, spp=spp, sppe=sppe, sppse=sppse))
    for leaf in leaves:
        if leaf[0] == "albedo":
            t = np.asarray([[1.0, 0.5, 0.25]], np.float32)
            r.set_bsdf_tangent(leaf[1], "reflectance", t); o.set_bsdf_tangent(leaf[1], "reflectance", t.reshape(1, 1, 3))
        elif leaf[0] == "vertices":
            t = rng.normal(size=(r.num_vertices(leaf[1]), 3)).astype(np.float32)
            r.set_mesh_vertex_tangent(leaf[1], t); o.set_mesh_vertex_tangent(leaf[1], t)
        elif leaf[0] == "rc":
            t = np.asarray([[0.7]], np.float32)
            r.set_bsdf_tangent(leaf[1], "alpha_u", t); o.set_bsdf_tangent(leaf[1], "alpha_u", t.reshape(1, 1, 1))
            t3 = rng.normal(size=(1, 3)).astype(np.float32)
            r.set_bsdf_tangent(leaf[1], "eta", t3); o.set_bsdf_tangent(leaf[1], "eta", t3.reshape(1, 1, 3))
    r.configure(); o.configure()
    ri, oi = refrun.DirectIntegrator(bs, ls), orc.DirectIntegrator(bs, ls)
    if guiding:
        ri.preprocess_secondary_edges(r, 0, guiding[0], guiding[1]); oi.preprocess_secondary_edges(o, 0, guiding[0], guiding[1])
    t0 = time.time(); ac = ri.renderC(r) if spp > 0 else None; (a, at) = ri.renderD(r); tr = time.time() - t0
    t0 = time.time(); bc = oi.renderC(o) if spp > 0 else None; (b, bt) = oi.renderD(o); to = time.time() - t0
    print("== %s: %s %dx%d spp %d/%d/%d Direct(%d,%d) leaves %s%s   [reference source %.1f s, oracle %.1f s]" % (
        label, name, w, h, spp, sppe, sppse, bs, ls, leaves, " guided %s x%d" % guiding if guiding else "", tr, to))
    if ac is not None:
        err = np.abs(ac - bc).mean(axis=1)
        print("   renderC : per-pixel L1 > 1e-4 on %d of %d px (max %.3g); image mean rel diff %.2e" % ((err > 1e-4).sum(), len(err), err.max(), abs(ac.mean() - bc.mean()) / bc.mean()))
        print("   renderD : " + stats(a, b, 2e-4))
    print("   d image : " + stats(at, bt, 1e-3))
    sys.stdout.flush()


def conditioning():
    """the reference's own code against itself with matrix * vector rounded differently in the last bit (fmadd chain vs plain sums): how much
    of the primary-edge term's difference to the oracle is fp32 conditioning of the estimator (ray pairs 1e-5 off a silhouette seen from
    1000 units away), not a difference in the algorithm"""
    xml = os.path.join(T, "data", "scenes", "cbox_bunny.xml")
    out = []
    for plain in (True, False):
        refrun.set_matvec_plain(plain)
        r = refrun.Scene(xml, T, 64, 64, 0, 16, 0)
        r.set_mesh_vertex_tangent(1, np.random.default_rng(17).normal(size=(r.num_vertices(1), 3)).astype(np.float32))
        r.configure()
        out.append(refrun.DirectIntegrator(1, 1).renderD(r)[1])
    refrun.set_matvec_plain(True)
    print("== conditioning: reference source vs reference source, primary edges only, cbox_bunny 64x64 sppe 16, plain vs fmadd matrix product")
    print("   d image : " + stats(out[0], out[1], 1e-3))


if __name__ == "__main__":
    if "--conditioning" in sys.argv:
        conditioning()
        sys.exit(0)
    run("cfg1 (full size)", "cbox_bunny", (128, 128, 16, 0, 0), [("albedo", 0)])
    run("cfg2 (reduced)", "cbox_bunny", (256, 256, 16, 0, 0), [("albedo", 0)])
    run("cfg3 (reduced)", "cbox_bunny", (128, 128, 8, 8, 8), [("vertices", 1)])
    run("cfg3 guided (reduced)", "cbox_bunny", (128, 128, 8, 8, 8), [("vertices", 1)], guiding=([40, 5, 5, 2], 4))
    run("cfg5 (reduced)", "bunny_env", (128, 128, 8, 8, 8), [("rc", 0), ("vertices", 0)])
    conditioning()
