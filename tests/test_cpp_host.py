"""The C++ host layer (cpp/feriphys_cuda.hpp) mirrors the reference's Rust API; its test
program replays the reference's own State tests (state.rs:166-280) and the demo's
Simulation call sequence through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "cpp", "_build", "test_host")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "cpp")], check=True, stdout=subprocess.DEVNULL)


def test_cpp_host_logic_without_gpu():
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "CPP_HOST_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_layer_on_gpu():
    _build()
    r = subprocess.run([BIN, "--gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CPP_GPU_OK" in r.stdout, r.stdout + r.stderr
