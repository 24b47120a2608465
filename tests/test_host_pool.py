"""The host thread pool of svs_ba_set_problem (scavislam_b200/csrc/host_pool.hpp): a stress test in plain C++,
built here with g++ (no GPU)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_spin_pool_runs_every_chunk_exactly_once(tmp_path):
    exe = str(tmp_path / "host_pool_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "host_pool_main.cpp")],
                   check=True)
    for _ in range(3):
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and out.stdout.strip() == "OK"
