"""GPU test: psdr-cuda's OWN example scripts, copied verbatim into tests/data/examples/ (zero edits), run against this repo's
`import psdr_cuda` / `import enoki` — BASELINE.json north_star "examples/ run unchanged". Covers examples/psdr_test.py (run_orig /
run_ad / run_fd through examples/run_test.py and examples/utils/differential.py: forward mode, `ek.forward`) and an optimisation
loop with examples/utils/adam.py (reverse mode: `ek.set_requires_gradient(bsdf.reflectance.data)`, `ek.backward(loss)`,
`ek.gradient`, docs/inverse_diff_render.rst:63-79). The driver only lowers `config.psdr_tests[...]` pass counts at run time."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
EX = os.path.join(ROOT, "tests", "data", "examples")


def _workdir(tmp_path):
    for f in ("psdr_test.py", "run_test.py", "config.py", "utils"):
        os.symlink(os.path.join(EX, f), tmp_path / f)
    os.symlink(os.path.join(ROOT, "tests", "data"), tmp_path / "data")      # ./data/scenes/*.xml, ./data/objects/...
    return tmp_path


def _run(tmp_path, body, timeout=900):
    driver = "import os, sys, json\nsys.path.insert(0, %r)\nos.environ['OPENCV_IO_ENABLE_OPENEXR'] = '1'\nimport psdr_cuda_b200.compat\nsys.path.insert(0, '.')\n" % ROOT
    out = subprocess.run([sys.executable, "-c", driver + textwrap.dedent(body)], cwd=str(tmp_path), capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("name,files,min_corr", [
    ("bunny_silhouette", ["field_AD_0.exr", "field_FD_0.exr"], 0.8),                       # field integrator, mesh_rotate, primary edges
    ("cbox_mutie", ["direct_orig_0.exr"], None),                                           # several emitters, renderC only
    ("bunny_env_1", ["direct_orig_0.exr", "direct_AD_0.exr", "direct_FD_0.exr"], 0.5),     # rough conductor + envmap, envmap_rotate
    ("cbox_MIS", ["direct_orig_0.exr", "direct_AD_0.exr", "direct_FD_0.exr"], None),       # vertex_transform, guided secondary edges
])
def test_psdr_test_py_runs_unchanged(native_lib, tmp_path, name, files, min_corr):
    wd = _workdir(tmp_path)
    r = _run(wd, """
        import numpy as np, cv2
        import config
        t = config.psdr_tests[%r]
        # fewer passes / guiding rounds than the shipped configuration: the scripts themselves are untouched
        t["npass"] = 2
        for k in ("AD", "FD"):
            if k in t:
                t[k] = dict(t[k]); t[k]["npass"] = 2 if k == "AD" else 4
        if "AD" in t and "guide" in t["AD"]:
            t["AD"]["guide"] = {"reso": [2000, 4, 4, 2], "nround": 2}
        import psdr_test
        os.makedirs(config.output_path, exist_ok=True)
        psdr_test.process(%r, t)
        d = config.output_path + t["fname"] + "/"
        res = {"files": sorted(os.listdir(d))}
        imgs = {f: cv2.imread(d + f, cv2.IMREAD_UNCHANGED) for f in res["files"]}
        res["finite"] = bool(all(np.isfinite(v).all() for v in imgs.values()))
        res["nonzero"] = {f: float(np.abs(v).max()) for f, v in imgs.items()}
        ad = [v for f, v in imgs.items() if "_AD_" in f]; fd = [v for f, v in imgs.items() if "_FD_" in f]
        if ad and fd:
            b = 32
            def blocks(x):
                h, w = x.shape[0] // b * b, x.shape[1] // b * b
                return x[:h, :w].reshape(h // b, b, w // b, b, 3).mean(axis=(1, 3, 4)).ravel()
            res["corr"] = float(np.corrcoef(blocks(ad[0]), blocks(fd[0]))[0, 1])
        print(json.dumps(res))
        """ % (name, name))
    for f in files:
        assert f in r["files"], r
    assert r["finite"]
    assert all(r["nonzero"][f] > 0 for f in files), r
    if min_corr is not None:
        assert r["corr"] >= min_corr, r          # AD and FD derivative images agree block-wise (Monte Carlo noise on both)


def test_adam_py_optimises_an_albedo_and_a_vertex_through_ek_backward(native_lib, tmp_path):
    wd = _workdir(tmp_path)
    r = _run(wd, """
        import numpy as np
        import psdr_cuda
        import enoki as ek
        from enoki.cuda_autodiff import Float32 as FloatD, Vector3f as Vector3fD
        from utils.adam import Adam

        def make():
            sc = psdr_cuda.Scene()
            sc.load_file("./data/scenes/cbox_bunny.xml", False)
            sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse, sc.opts.log_level = 64, 64, 256, 0, 0, 0
            return sc
        integrator = psdr_cuda.DirectIntegrator(bsdf_samples=1, light_samples=1)
        ref = make(); ref.configure()
        target = integrator.renderC(ref, 0)                               # the shipped albedo is the optimum
        sc = make()
        pm = sc.param_map
        pm["BSDF[0]"].reflectance.data = Vector3fD(np.array([[0.4, 0.6, 0.3]], np.float32))    # start somewhere else
        opt = Adam({0: pm["BSDF[0]"]}, {}, [0], [], lr=0.05)
        losses, albedos = [], []
        for it in range(40):
            sc.configure()
            img = integrator.renderD(sc, 0)
            loss = ek.hmean(ek.squared_norm(img - target))
            ek.backward(loss)
            g = ek.gradient(pm["BSDF[0]"].reflectance.data).numpy()
            opt.step()
            losses.append(float(loss.numpy()[0])); albedos.append(pm["BSDF[0]"].reflectance.data.numpy()[0].tolist())
        # docs/inverse_diff_render.rst:63-79 verbatim flow on the vertex positions: gradient exists, is finite, and moves the loss
        sc2 = make()
        sc2.opts.spp, sc2.opts.sppe, sc2.opts.sppse = 16, 8, 8
        ek.set_requires_gradient(sc2.param_map["Mesh[1]"].vertex_positions)
        sc2.configure()
        image = psdr_cuda.DirectIntegrator().renderD(sc2, sensor_id=0)
        loss2 = ek.sqrt(ek.hmean(ek.squared_norm(target * 0.9 - image)))
        ek.backward(loss2)
        grad = ek.gradient(sc2.param_map["Mesh[1]"].vertex_positions).numpy()
        print(json.dumps({"losses": losses, "albedo": albedos[-1], "albedo0": albedos[0], "g_last": g.tolist(), "vgrad_shape": list(grad.shape),
                          "vgrad_finite": bool(np.isfinite(grad).all()), "vgrad_norm": float(np.linalg.norm(grad))}))
        """)
    # every pass draws new samples, so the loss keeps a Monte Carlo floor (2 x the per-pixel variance); the albedo is the criterion
    assert r["losses"][-1] < 0.5 * r["losses"][1], r["losses"]
    assert max(abs(a - 0.95) for a in r["albedo"]) < 0.1, (r["albedo0"], r["albedo"])     # back at the white albedo of cbox_bunny.xml (0.95)
    assert r["vgrad_shape"][1] == 3 and r["vgrad_finite"] and r["vgrad_norm"] > 0
