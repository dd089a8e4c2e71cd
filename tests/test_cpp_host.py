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


def _mix(x):
    m = (1 << 64) - 1
    z = (x + 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


@pytest.mark.gpu
def test_headless_demo_driver_matches_the_python_mirror():
    """cpp/demo_flocking.cpp replays demos/flocking.rs:92-156 + update() (:209-231) headless; the
    Python mirror running the same scene and frame loop must upload bit-identical instances."""
    import numpy as np

    from feriphys_b200.flocking import demo_simulation
    _build()
    frames, frame_ms = 4, 16
    r = subprocess.run([os.path.join(ROOT, "cpp", "_build", "demo_flocking"), str(frames), str(frame_ms)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    fields = r.stdout.split()
    got = dict(zip(fields[0::2], fields[1::2]))
    sims = [demo_simulation(1), demo_simulation(2)]
    h = 0
    for _ in range(frames):
        for s in sims:
            for _ in range(frame_ms):       # dt = 1 ms: frame_ms steps per frame
                s.step()
        for s in sims:
            for w in s.read_instances().view(np.uint32).reshape(-1):
                h = _mix(h ^ int(w))
    assert int(got["steps"]) == 2 * frames * frame_ms
    assert got["checksum"] == f"{h:016x}", (got, f"{h:016x}")
