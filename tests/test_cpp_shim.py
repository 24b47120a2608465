"""The C++ host layer (include/svs_b200.hpp): compiles with plain g++ against the C ABI (CPU), and on
the GPU box reproduces the oracle through the id-based SlamGraph-like interface."""
import os
import subprocess

import numpy as np
import pytest

from scavislam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "shim_main")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "shim_main.cpp")
    lib_dir = os.path.join(ROOT, "scavislam_b200")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "include", "svs_b200.hpp"))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L", lib_dir, "-lsvsb200", f"-Wl,-rpath,{lib_dir}"])
    return EXE


def _dump(pb, path):
    xyz = np.stack([pb.psi[:, 0] / pb.psi[:, 2], pb.psi[:, 1] / pb.psi[:, 2], 1.0 / pb.psi[:, 2]], 1)
    with open(path, "wb") as f:
        np.array([pb.P, pb.L, pb.E, pb.C], np.int32).tofile(f)
        np.asarray(pb.cam, np.float64).tofile(f)
        for a in (pb.pose_qt, xyz):
            np.ascontiguousarray(a, np.float64).tofile(f)
        for a in (pb.e_point, pb.e_pose, pb.e_anchor):
            np.ascontiguousarray(a, np.int32).tofile(f)
        for a in (pb.e_obs, pb.e_info):
            np.ascontiguousarray(a, np.float64).tofile(f)
        for a in (pb.c_i, pb.c_j):
            np.ascontiguousarray(a, np.int32).tofile(f)
        for a in (pb.c_T, pb.c_Lambda):
            np.ascontiguousarray(a, np.float64).tofile(f)


def test_cpp_layer_compiles_and_fails_loudly_without_gpu(svs, tmp_path):
    import torch
    exe = _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    pb = synth.make_window(4, 30, seed=2)
    _dump(pb, tmp_path / "in.bin")
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), "1"], capture_output=True, text=True)
    assert r.returncode == 3 and "NO_GPU" in r.stdout        # no CPU fallback


@pytest.mark.gpu
def test_cpp_layer_matches_oracle(svs, oracle, tmp_path):
    exe = _build()
    pb = synth.make_config("C1")
    _dump(pb, tmp_path / "in.bin")
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK iterations=2" in r.stdout and f"point_edges={pb.E}" in r.stdout
    out = np.fromfile(tmp_path / "out.bin", np.float64)
    poses = out[:7 * pb.P].reshape(pb.P, 7)
    xyz = out[7 * pb.P:].reshape(pb.L, 3)
    p_o, s_o, _ = oracle.optimize(pb, 2)     # huber delta stays 1 like the reference
    xyz_o = np.stack([s_o[:, 0] / s_o[:, 2], s_o[:, 1] / s_o[:, 2], 1.0 / s_o[:, 2]], 1)
    assert np.abs(poses - p_o).max() <= 1e-6 * np.abs(p_o).max()
    assert np.abs(xyz - xyz_o).max() <= 1e-6 * np.abs(xyz_o).max()
