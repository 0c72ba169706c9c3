"""GPU test: the reference examples' validation flow (AD derivative image via ek.forward vs central finite differences,
examples/run_test.py:44-231) runs against `import psdr_cuda` / `import enoki` of this repo, in a fresh interpreter."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scene,kind,min_corr", [("cbox_bunny.xml", "mesh_transform", 0.8), ("bunny.xml", "mesh_rotate", 0.8), ("cbox_bunny.xml", "vertex_transform", 0.6)])
def test_ad_matches_fd_through_the_reference_surface(native_lib, scene, kind, min_corr):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "examples_flow", "run_flow.py"), scene, kind], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["finite"] and r["nonzero"] > 0
    assert r["corr"] >= min_corr, r                      # AD and FD derivative images agree block-wise (Monte Carlo noise on both)
    assert r["sum_ad"] * r["sum_fd"] > 0 or abs(r["sum_fd"]) < 0.05 * abs(r["sum_ad"]) + 1e-3, r
