"""The C++ mirror of the crate API (rust-la_b200/host/la.hpp) replays the reference's unit tests in a compiled language
(host/test_la.cpp, built by `make -C rust-la_b200 all` / __graft_entry__.build()).  Run the binary on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "rust-la_b200", "build", "test_la")


@pytest.mark.gpu
def test_cpp_mirror_replays_reference_tests():
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "rust-la_b200"), "build/test_la"])
    out = subprocess.run([BIN, "--require-gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failed" in out.stdout and "FAILED" not in out.stdout, out.stdout
