"""BASELINE.json's configurations at their full sizes, checked through properties that do
not need the (hours-long) CPU oracle: injected integer delays are recovered in every
window and pair, the FFT path agrees with the exhaustive time-domain path, closure of the
lags around station triangles, shift covariance, and the grid arg-min against the oracle
on the neighbourhood of the winning cell."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

import tdoa_b200 as T
from oracle import oracle

pytestmark = pytest.mark.gpu

FS, C = 2e6, 299792458.0


@pytest.fixture(scope="module")
def gen():
    import torch
    import bench
    import configs_bench as cb
    return torch, bench, cb, torch.device("cuda", 0)


def load_dev(e, caps):
    for k, c in enumerate(caps):
        e.load_u8_device(k, c.data_ptr(), c.numel(), keep=c)


def test_config2_full_length_fft_equals_exhaustive(gen):
    """configs[1]/[2]: 3 x 200e6-sample captures, whole-signal correlation."""
    torch, bench, cb, dev = gen
    delays, _ = cb.delays_for(bench.STATION_LLH)
    caps = cb.synth(dev, 3, 66_666_666, delays)
    want = [int(delays[j] - delays[i]) for i in range(3) for j in range(i + 1, 3)]
    out = {}
    for use_fft in (1, 0):
        with T.Engine(T.MODE_BINARY, chunk_samples=0, use_fft=use_fft) as e:
            load_dev(e, caps)
            out[use_fft] = e.xcorr(T.KIND_TGT)[0]
            if use_fft:
                ref = e.xcorr(T.KIND_REF)[0]
    assert [int(x) for x in out[1]["lag"]] == want == [int(x) for x in out[0]["lag"]]
    assert [int(x) for x in ref["lag"]] == want
    assert np.max(np.abs(out[1]["corr"] - out[0]["corr"])) <= 1e-12
    assert np.all(out[1]["n_blocks"] == 6666) and np.all(ref["n_blocks"] == 13333)
    assert np.all((out[1]["flags"] >> 16 & 255) <= 4)  # a sharp peak needs few exact evaluations


def test_shift_covariance(gen):
    """Dropping k samples from the front of one station's target block moves its lags by k."""
    torch, bench, cb, dev = gen
    delays, _ = cb.delays_for(bench.STATION_LLH)
    block = 3_000_000
    caps = cb.synth(dev, 3, block, delays)
    with T.Engine(T.MODE_BINARY) as e:
        load_dev(e, caps)
        base = e.xcorr(T.KIND_TGT)[0]["lag"].astype(int)
        k = 17
        shifted = caps[2].clone()
        shifted[2 * block + 2 * k:4 * block] = caps[2][2 * block:4 * block - 2 * k]  # delay station 2 by k more
        e.load_u8_device(2, shifted.data_ptr(), shifted.numel(), keep=shifted)
        moved = e.xcorr(T.KIND_TGT)[0]["lag"].astype(int)
    assert list(moved - base) == [0, k, k]


def test_config3_every_window_recovers_the_delays(gen):
    torch, bench, cb, dev = gen
    delays, _ = cb.delays_for(bench.STATION_LLH)
    block, W, L = 66_666_666, 2_000_000, 50_000
    caps = cb.synth(dev, 3, block, delays)
    want = [int(delays[j] - delays[i]) for i in range(3) for j in range(i + 1, 3)]
    nw = block // W
    with T.Engine(T.MODE_EXTENDED, max_lag=L) as e:
        load_dev(e, caps)
        pk = e.xcorr(T.KIND_TGT, 0, W, nw, W)
        rev = e.xcorr(T.KIND_TGT, 5 * W, W, 1, 0)[0]
    assert pk.shape == (33, 3)
    for w in range(nw):
        assert [int(x) for x in pk[w]["lag"]] == want
    assert np.all(np.abs(pk["frac"]) < 0.25) and np.all(pk["corr"] > 0.5)
    assert np.array_equal(rev["lag"], pk[5]["lag"]) and np.array_equal(rev["corr"], pk[5]["corr"])


def test_wide_lags_long_windows_segmented_big_transform(gen):
    """Templates longer than one 2^21-point transform: the cross-spectra of the segments are
    added before the one inverse.  Same records as the 2048-lag chunk path (use_fft=3)."""
    torch, bench, cb, dev = gen
    delays, _ = cb.delays_for(bench.STATION_LLH)
    block, W, L = 12_000_000, 5_000_000, 6_000
    caps = cb.synth(dev, 3, block, delays)
    want = [int(delays[j] - delays[i]) for i in range(3) for j in range(i + 1, 3)]
    out = []
    for use_fft in (1, 3):
        with T.Engine(T.MODE_EXTENDED, max_lag=L, use_fft=use_fft) as e:
            load_dev(e, caps)
            out.append(e.xcorr(T.KIND_TGT, 1000, W, 2, W))
    for w in range(2):
        assert [int(x) for x in out[0][w]["lag"]] == want
    for name in ("lag", "corr", "frac"):
        assert np.array_equal(out[0][name], out[1][name]), name


def test_config4_sixteen_stations_closure(gen):
    torch, bench, cb, dev = gen
    st = cb.ring_stations(16)
    d16, _ = cb.delays_for(st)
    W = 2_000_000
    caps = cb.synth(dev, 16, 2 * W, d16)
    with T.Engine(T.MODE_EXTENDED, n_stations=16, max_lag=2000) as e:
        load_dev(e, caps)
        pk = e.xcorr(T.KIND_TGT, 0, W, 2, W)
    pairs = [(i, j) for i in range(16) for j in range(i + 1, 16)]
    for w in range(2):
        lag = {p: int(x) for p, x in zip(pairs, pk[w]["lag"])}
        assert all(lag[(i, j)] == int(d16[j] - d16[i]) for i, j in pairs)
        assert all(lag[(i, j)] + lag[(j, k)] == lag[(i, k)] for i in range(16) for j in range(i + 1, 16)
                   for k in range(j + 1, 16))


def test_config5_grid_million_cells(gen):
    torch, bench, cb, dev = gen
    st = cb.ring_stations(16)
    _, dist = cb.delays_for(st)
    rd = np.array([dist[j] - dist[i] for i in range(16) for j in range(i + 1, 16)])
    rng = np.random.default_rng(5)
    rds = rd[None, :] + rng.normal(0, 50e-9 * C, (16, rd.size))
    desc = [41.26 - 0.25, -96.02 - 0.25, 0.0005, 0.0005, 1000, 1000, 400.0]
    with T.Engine(T.MODE_BINARY) as e:
        out, cost, idx = e.grid(st, desc, rds)
    for k in (0, 7, 15):
        a, b = divmod(int(idx[k]), 1000)
        a0, b0 = max(a - 5, 0), max(b - 5, 0)
        wi, wc, _ = oracle.grid_solve(st, rds[k], desc[0] + a0 * desc[2], desc[1] + b0 * desc[3], desc[2], desc[3],
                                      11, 11, desc[6])
        assert (a0 + wi // 11, b0 + wi % 11) == (a, b)       # the oracle finds the same cell locally
        assert cost[k] == pytest.approx(wc, rel=1e-9)
        err = np.linalg.norm(bench.llh_to_ecef(*out[k]) - bench.llh_to_ecef(*bench.TX_LLH))
        assert err < 120.0                                   # within two cells of the transmitter
