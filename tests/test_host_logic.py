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
    # encoding/csv rejects a short record before loadStations' own check can (what the binary prints)
    with pytest.raises(RuntimeError, match="failed to read CSV: record on line 3: wrong number of fields"):
        T.TDOAProcessor(162400000.0, 92300000.0, str(bad), out=io.StringIO())
    bad.write_text("Name,Latitude,Longitude\n162400000,1,2\n")
    with pytest.raises(RuntimeError, match="invalid CSV format at line 2"):   # processor.go:67-69
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


def _cli_golden():
    import json
    from pathlib import Path
    return json.loads((Path(__file__).resolve().parent / "golden" / "cli_errors.json").read_text())


def _cli_setup(tmp_path):
    from pathlib import Path
    gold = _cli_golden()
    for name, text in gold["files"].items():
        (tmp_path / name).write_bytes(text.encode())
    csv_path = str(Path(__file__).resolve().parent / "golden" / "stations.csv")
    return gold, csv_path


CLI_CASE_NAMES = sorted(set(_cli_golden()["cases"]) - set(_cli_golden()["engine_cases"]))   # the rest needs the engine: test_gpu_parity.py


@pytest.mark.parametrize("case", CLI_CASE_NAMES)
def test_cli_python_mirror_matches_the_reference_binary(tmp_path, capsys, case):
    """Usage, argument and station-table errors: stdout, stderr (after the log time stamp) and
    exit status of the reference's own binary (tests/golden/cli_errors.json, processor.go:1047-1076,
    :52-107, :740-742) -- everything the command does before it touches a sample."""
    import re
    from importlib import import_module
    proc = import_module("tdoa-geolocation_b200.processor")
    gold, csv_path = _cli_setup(tmp_path)
    want = gold["cases"][case]
    argv = [a.replace("{csv}", csv_path).replace("{dir}", str(tmp_path)) for a in want["args"]]
    rc = proc.main(argv, prog="{prog}")
    got = capsys.readouterr()
    back = lambda t: t.replace(csv_path, "{csv}").replace(str(tmp_path), "{dir}")
    assert rc == want["returncode"]
    assert [back(l) for l in got.out.splitlines()] == want["stdout"]
    err = got.err.splitlines()
    assert all(re.match(r"\d{4}/\d\d/\d\d \d\d:\d\d:\d\d ", l) for l in err)
    assert [back(l[20:]) for l in err] == want["stderr"]


@pytest.mark.parametrize("case", CLI_CASE_NAMES)
def test_cli_cpp_mirror_matches_the_reference_binary(tmp_path, case):
    """The same, for processor_b200 (host/processor_b200.cpp)."""
    import re
    import subprocess
    exe = str(_host_binary())
    gold, csv_path = _cli_setup(tmp_path)
    want = gold["cases"][case]
    argv = [a.replace("{csv}", csv_path).replace("{dir}", str(tmp_path)) for a in want["args"]]
    r = subprocess.run([exe, *argv], capture_output=True, text=True, cwd=tmp_path)
    back = lambda t: t.replace(csv_path, "{csv}").replace(str(tmp_path), "{dir}").replace(exe, "{prog}")
    assert r.returncode == want["returncode"]
    assert [back(l) for l in r.stdout.splitlines()] == want["stdout"]
    err = r.stderr.splitlines()
    assert all(re.match(r"\d{4}/\d\d/\d\d \d\d:\d\d:\d\d ", l) for l in err)
    assert [back(l[20:]) for l in err] == want["stderr"]


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
    assert r.returncode == 1 and r.stdout.startswith(f"Usage: {exe} <ref_freq_hz> <target_freq_hz> <csv_file> <dat_file1>")
    golden = str(_host_binary().parent.parent / "tests" / "golden" / "stations.csv")
    r = subprocess.run([str(exe), "1", "92300000", golden, "a.dat", "b.dat", "c.dat"], capture_output=True, text=True)
    assert r.returncode == 1 and "reference frequency 1 not found in stations" in r.stderr
    r = subprocess.run([str(exe), "162400000", "92300000", golden, "a.dat", "b.dat"], capture_output=True, text=True)
    assert r.returncode == 1 and "need at least 3 collector stations, got 2" in r.stderr   # processor.go:740-742


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
