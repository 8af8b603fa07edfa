"""CPU check of the discriminator's arctangent (csrc/atan2_core.cuh) against libm."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host compile only)")
def test_atan2_core_matches_libm(tmp_path):
    exe = tmp_path / "atan2_check"
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I",
                    str(ROOT / "tdoa-geolocation_b200" / "csrc"), "-o", str(exe),
                    str(ROOT / "tests" / "native" / "atan2_check.cu")], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
