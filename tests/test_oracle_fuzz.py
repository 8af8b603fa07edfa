"""The CPU oracle against the reference's own shipped binary on random capture sets (CPU only: the
binary is run here, oracle/_ref/processor).  Complements the committed golden vectors: every record
(delay bit-exact, correlation to the printed six decimals), every preprocessing branch the binary
announces, and the outcome / trace of its solver."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle
from helpers import GOLDEN, STATION_LLH, STATIONS
from test_oracle_golden import check_binary_solver_against_stdout

REF_BINARY = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "processor"

_spec = importlib.util.spec_from_file_location("fuzz_binary", GOLDEN / "fuzz_binary.py")
fuzz = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(fuzz)


@pytest.mark.skipif(not REF_BINARY.exists(), reason="oracle/_ref/processor is staged by oracle/Makefile where /root/reference exists")
@pytest.mark.parametrize("seed", [1, 6, 9, 21, 23, 30, 100, 102, 119])   # 100: ten solver iterations; 102 / 119: its solver converges after 9 / 7
def test_oracle_equals_the_binary_on_random_captures(seed):
    caps = fuzz.make(seed)
    parsed, text, rc, err = fuzz.run_binary(caps, REF_BINARY, GOLDEN / "stations.csv")
    raws = [caps[s] for s in STATIONS]
    ref, tgt = oracle.process_capture_binary(raws)
    got = [("REF",) + r for r in ref] + [("TGT",) + r for r in tgt]
    assert len(parsed["pairs"]) == 6
    for (kind, delay, corr, _), want in zip(got, parsed["pairs"]):
        assert (kind, delay) == (want["kind"], want["delay"]), (seed, want)
        assert abs(corr - want["corr"]) <= 0.51e-6, (seed, corr, want)
    # the branch of every preprocessSignal call, in the binary's call order
    sigs = []
    for kind in ("ref", "tgt"):
        per = []
        for raw in raws:
            d = oracle.unpack_u8(raw)
            per.append((oracle.extract_reference(d) if kind == "ref" else oracle.extract_target(d))[:1_000_000])
        for i in range(3):
            for j in range(i + 1, 3):
                sigs += [per[i], per[j]]
    assert [oracle.preprocess_binary(s)[1] for s in sigs] == parsed["branch"]
    rd = [(t["delay"] / 2e6 - r["delay"] / 2e6) * 299792458.0 for r, t in zip(parsed["pairs"][:3], parsed["pairs"][3:])]
    check_binary_solver_against_stdout(STATION_LLH, rd, text, rc, err)
