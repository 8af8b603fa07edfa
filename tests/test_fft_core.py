"""CPU check of the shared-memory FFT's per-thread phases (csrc/fft_core.cuh): the
host emulation in tests/native/fft_emul.cu runs the same __host__ __device__ code one
"thread" at a time and compares the 8192-point result with a direct DFT."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host compile only)")
def test_fft_phases_match_direct_dft(tmp_path):
    exe = tmp_path / "fft_emul"
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I",
                    str(ROOT / "tdoa-geolocation_b200" / "csrc"), "-o", str(exe),
                    str(ROOT / "tests" / "native" / "fft_emul.cu")], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    rel = float(res.stdout.split()[1])
    assert rel < 1e-6


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host compile only)")
def test_tile_fft_phases_and_cross_spectra(tmp_path):
    """csrc/fft_tile_core.cuh (the tile kernel's transform: DIT-FMA butterflies, adjacent
    butterflies per thread, 128-bit row layout) and the 2 x 2 tile cross-spectrum formulas,
    emulated thread by thread on the host against direct DFTs in double precision."""
    exe = tmp_path / "fft_tile_emul"
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I",
                    str(ROOT / "tdoa-geolocation_b200" / "csrc"), "-o", str(exe),
                    str(ROOT / "tests" / "native" / "fft_tile_emul.cu")], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    vals = {l.split()[0]: float(l.split()[1]) for l in res.stdout.splitlines()}
    assert vals["rel_rms_err"] < 1e-6 and vals["cross_rel_rms_err"] < 2e-6


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host compile only)")
def test_tile16_fft_phases_match_direct_dft(tmp_path):
    """csrc/fft_tile16_core.cuh (512 threads x 16 points: 16 x 16 x 32 with the radix-32 step taken by pairs of
    lanes through a shuffle), emulated thread by thread on the host against a direct DFT in double precision."""
    exe = tmp_path / "fft_tile16_emul"
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I",
                    str(ROOT / "tdoa-geolocation_b200" / "csrc"), "-o", str(exe),
                    str(ROOT / "tests" / "native" / "fft_tile16_emul.cu")], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    assert float(res.stdout.split()[1]) < 1e-6


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host compile only)")
def test_chunk_parallel_sequential_sum_is_exact(tmp_path):
    """csrc/seqsum_core.cuh: the reference's sequential f32 accumulator (processor.go:304-309) evaluated
    chunk by chunk in integer arithmetic, against the plain loop, bit for bit, on 49 signals (noise,
    DC-heavy, a tie at every step, sign changes, 8 decades of dynamic range, tiny inputs)."""
    exe = tmp_path / "seqsum_emul"
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I",
                    str(ROOT / "tdoa-geolocation_b200" / "csrc"), "-o", str(exe),
                    str(ROOT / "tests" / "native" / "seqsum_emul.cu")], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    last = res.stdout.strip().splitlines()[-1].split()
    assert last[0] == "mismatch" and int(last[1]) == 0
    assert float(last[3]) > 0.9      # DC-heavy signals: the O(1) path carries almost every chunk
