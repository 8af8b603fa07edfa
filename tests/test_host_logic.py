"""Host-side mirror of the reference's interface (tdoa-geolocation_b200/processor.py):
station table, filename->station, CLI and error behaviour (processor.go:36-122, 1047-1076)."""
import io

import pytest

import tdoa_b200 as T
from helpers import GOLDEN

CSV = GOLDEN / "stations.csv"


def make(ref=162400000.0):
    return T.TDOAProcessor(ref, 92300000.0, str(CSV), out=io.StringIO())


def test_load_stations_and_reference_row():
    p = make()
    assert len(p.stations) == 5  # processor.go:66 skips the header
    assert p.ref_station.name == "162400000"  # :96  "%.0f" of the reference frequency
    assert abs(p.stations["kx0u"].latitude - 41.18660274289527) < 1e-12
    assert "Loaded 5 stations including reference 162 MHz" in p.out.getvalue()


def test_missing_reference_frequency_is_an_error():
    with pytest.raises(RuntimeError, match="reference frequency 100000000 not found"):
        make(ref=100e6)


def test_missing_csv_is_an_error(tmp_path):
    with pytest.raises(RuntimeError, match="failed to open CSV"):
        T.TDOAProcessor(162400000.0, 92300000.0, str(tmp_path / "nope.csv"), out=io.StringIO())


def test_bad_csv_row(tmp_path):
    bad = tmp_path / "bad.csv"
    bad.write_text("Name,Latitude,Longitude,Elevation\n162400000,1,2,3\nx,1,2\n")
    with pytest.raises(RuntimeError, match="invalid CSV format at line 3"):
        T.TDOAProcessor(162400000.0, 92300000.0, str(bad), out=io.StringIO())


def test_station_from_filename():
    p = make()
    assert p.get_station_from_filename("/data/sim-kx0u-1754900000.dat").name == "kx0u"
    assert p.get_station_from_filename("weak-kf0mtl-17.dat").name == "kf0mtl"
    with pytest.raises(RuntimeError, match="could not identify station"):
        p.get_station_from_filename("nobody-1.dat")


def test_process_needs_three_collectors():
    p = make()
    with pytest.raises(RuntimeError, match="need at least 3 collector stations, got 2"):
        p.process_tdoa(["a-kx0u.dat", "b-n3pay.dat"])


def test_cli_usage():
    from importlib import import_module
    proc = import_module("tdoa-geolocation_b200.processor")
    assert proc.main(["1", "2", "x.csv"]) == 1


# ------------------------------------------------------------------ the C++ host mirror (processor_b200)
def _host_binary():
    from pathlib import Path
    return Path(__file__).resolve().parent.parent / "tdoa-geolocation_b200" / "processor_b200"


def test_processor_b200_usage_and_station_table():
    """processor.go:1048-1051 usage line and exit status; loadStations (:52-107) runs on the host."""
    import subprocess
    exe = _host_binary()
    assert exe.exists(), "run python tdoa-geolocation_b200/build.py"
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("Usage: processor <reference_freq_hz> <target_freq_hz> <stations.csv>")
    golden = str(_host_binary().parent.parent / "tests" / "golden" / "stations.csv")
    r = subprocess.run([str(exe), "1", "92300000", golden, "a.dat", "b.dat", "c.dat"], capture_output=True, text=True)
    assert r.returncode == 1 and "reference frequency 1 not found in stations" in r.stderr
    r = subprocess.run([str(exe), "162400000", "92300000", golden, "a.dat", "b.dat"], capture_output=True, text=True)
    assert r.returncode == 1   # fewer than three collectors: the usage line (processor.go:1048)


def test_processor_b200_has_no_cpu_path():
    import subprocess
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("checks the no-GPU refusal")
    golden = str(_host_binary().parent.parent / "tests" / "golden" / "stations.csv")
    r = subprocess.run([str(_host_binary()), "162400000", "92300000", golden, "a.dat", "b.dat", "c.dat"],
                       capture_output=True, text=True)
    assert r.returncode == 1
    assert "Loaded 5 stations including reference 162 MHz" in r.stdout
    assert "TDOA processing failed" in r.stderr and "no CPU path" in r.stderr
