"""CPU test: the reverse-mode building blocks of csrc/pb_adjoint_math.cuh (normalize, ray/triangle, connection factor,
shading normal, face normal/area) against central finite differences. Host-compiled with nvcc; no GPU needed."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


def test_adjoint_helpers_match_finite_differences(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "adjoint_check")
    subprocess.check_call([nvcc, "-x", "cu", "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "adjoint_check.cu")], stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "adjoint_check: ok" in out.stdout, out.stdout[-2000:]


def test_roughconductor_dual_derivatives_match_finite_differences(tmp_path):
    """csrc/pb_rc.cuh: the local forward-mode duals of a rough-conductor event (geometry inputs and BSDF parameters, BSDF- and
    emitter-sampled connections, camera / path-space vertices, with and without MIS) against central differences."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "rc_dual_check")
    subprocess.check_call([nvcc, "-x", "cu", "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "rc_dual_check.cu")], stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "rc_dual_check: ok" in out.stdout, out.stdout[-2000:]


def test_device_lbvh_node_arithmetic_builds_a_valid_tree(tmp_path):
    """csrc/pb_lbvh.cuh (Morton codes, common-prefix lengths with the index tie-break, Karras range / split search, leaf collapse), the
    per-node arithmetic of the device LBVH build, host-compiled: 64 cases (17 .. 50 000 primitives; uniform, clustered, mostly-duplicate and
    seven-code inputs; 1 / 2 / 4 / 8 slots per leaf) must each give a binary tree that covers every sorted slot exactly once"""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "lbvh_check")
    subprocess.check_call([nvcc, "-x", "cu", "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "lbvh_check.cu")], stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "lbvh_check: ok" in out.stdout, out.stdout[-2000:]

