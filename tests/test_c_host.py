"""The drop-in boundary is a C ABI: ``include/d3roma_b200.h`` must compile as plain C99 and a C host must link against
``libd3roma_b200.so`` and run (examples/c_host.c: version, error protocol; with a GPU also one fused DDIM update)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")


def _build(tmp_path):
    from d3roma_b200 import _lib, build
    if not _lib.LIB_PATH.exists():
        build.build(verbose=False)
    exe = tmp_path / "c_host"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"),
           str(ROOT / "examples" / "c_host.c"), "-L", str(_lib.LIB_PATH.parent), "-ld3roma_b200",
           f"-Wl,-rpath,{_lib.LIB_PATH.parent}", "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_header_is_plain_c_and_c_host_runs_without_gpu(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "libd3roma_b200 version" in res.stdout and "last error text:" in res.stdout


@pytest.mark.gpu
def test_c_host_runs_a_kernel(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([str(exe), "gpu"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "d3_sched_ddim_step ran on the device" in res.stdout
