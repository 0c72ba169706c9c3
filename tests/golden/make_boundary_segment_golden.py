"""Golden vectors of Scene::sample_boundary_segment_direct (src/scene/scene.cpp:456-492) from the reference's OWN source
(oracle/_ref/libref_render.so, built by oracle/build_ref.sh from /root/reference): 4096 fixed samples on cbox_bunny and bunny_env.
    python tests/golden/make_boundary_segment_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402

out = {}
rng = np.random.default_rng(2024)
s3 = rng.uniform(0, 1, size=(4096, 3)).astype(np.float32)
out["sample3"] = s3
refrun.set_matvec_plain(True)
for name in ("cbox_bunny", "bunny_env"):
    sc = refrun.Scene(os.path.join(ROOT, "tests", "data", "scenes", name + ".xml"), os.path.join(ROOT, "tests"), 32, 32, 1, 1, 1)
    sc.configure()
    out[name] = sc.sample_boundary_segment_direct(s3)
    print(name, "valid", int(out[name][:, 16].sum()), "of", len(s3))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "boundary_segment_golden.npz"), **out)
