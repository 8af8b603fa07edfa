"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/tdoa_b200.h declares, has the documented record layout and defaults, and
refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import tdoa_b200 as T
from importlib import import_module

N = import_module("tdoa-geolocation_b200._native")
ROOT = Path(__file__).resolve().parent.parent


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def header_symbols():
    text = (ROOT / "include" / "tdoa_b200.h").read_text()
    return sorted(set(re.findall(r"TDOA_API\s+[\w\s\*]+?\b(tdoa_\w+)\s*\(", text)))


def test_library_is_in_tree_and_loads():
    assert T.library_path().exists(), "run python tdoa-geolocation_b200/build.py"
    assert ROOT in T.library_path().parents
    T.load_library()


def test_every_header_symbol_is_exported():
    lib = T.load_library()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tdoa_b200.h but not exported"
    assert sorted(N.ABI_SYMBOLS) == syms


def test_record_layouts():
    assert C.sizeof(N.PeakStruct) == 32 and N.PEAK_DTYPE.itemsize == 32
    assert C.sizeof(N.Config) == 8 + 9 * 4 + 7 * 4
    for name, _ in N.PeakStruct._fields_:
        assert N.PEAK_DTYPE.fields[name][1] == getattr(N.PeakStruct, name).offset


def test_default_configs_match_reference_constants():
    src = T.default_config(T.MODE_SOURCE)
    # processor.go:772 (chunk), :633 (maxLag), :682 (block), :821 (fs)
    assert (src.chunk_samples, src.max_lag, src.block_size, src.sample_rate) == (2000000, 20000, 1000, 2e6)
    b = T.default_config(T.MODE_BINARY)
    # shipped binary: 1 000 000-sample chunk, maxLag 2000, 10000-sample blocks, 120-sample sanity window
    assert (b.chunk_samples, b.max_lag, b.block_size, b.sanity_lag) == (1000000, 2000, 10000, 120)
    with pytest.raises(T.TdoaError):
        T.default_config(7)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU refusal")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(T.TdoaError) as ei:
        T.Engine(T.MODE_BINARY)
    assert ei.value.code == -2  # TDOA_E_NODEVICE
    assert "no CPU path" in str(ei.value)


def test_create_rejects_bad_config():
    lib = T.load_library()
    cfg = T.default_config(T.MODE_BINARY)
    cfg.n_stations = 1
    h = C.c_void_p()
    assert lib.tdoa_create(C.byref(h), C.byref(cfg)) == -1
    assert b"n_stations" in lib.tdoa_last_error(None)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = ROOT / "tdoa-geolocation_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        if "build" in p.parts:
            continue
        text = p.read_text()
        for line in text.splitlines():
            code = line.split("//")[0].split("#")[0]
            assert "import oracle" not in code and "from oracle" not in code, p
            assert "tdoa_oracle" not in code and "libtdoa_oracle" not in code, p
