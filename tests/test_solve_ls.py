"""Least-squares fix over all pair range differences.

The reference has no such solver (solveTDOA uses two of three measurements, freezes ECEF Z
and returns its 10th half-step, processor.go:932-1020), so parity is UNPINNED by the
reference: the oracle statement (orc_solve_ls) is checked against the geometry itself --
exact range differences of a known transmitter must give that transmitter back -- and the
GPU kernel against the oracle.
"""
import numpy as np
import pytest

from oracle import oracle
from helpers import STATION_LLH

C_LIGHT = 299792458.0
TX = np.array([41.20, -96.00, 400.0])  # simulator.go:229 usage example


def ring(n, seed=4242):
    rng = np.random.default_rng(seed)
    st = [list(s) for s in STATION_LLH]
    for k in range(n - 3):
        ang = 2 * np.pi * (k + rng.uniform(-0.2, 0.2)) / (n - 3)
        r_km = 25.0 * rng.uniform(0.8, 1.2)
        st.append([41.26 + r_km / 111.0 * np.cos(ang),
                   -96.02 + r_km / (111.0 * np.cos(np.radians(41.26))) * np.sin(ang), rng.uniform(300, 3000)])
    return np.array(st[:n])


def range_diffs(st, tx):
    x = oracle.llh_to_ecef(*tx)
    r = np.array([np.linalg.norm(x - oracle.llh_to_ecef(*s)) for s in st])
    return np.array([r[j] - r[i] for i in range(len(st)) for j in range(i + 1, len(st))])


def dist_m(a, b):
    return float(np.linalg.norm(oracle.llh_to_ecef(*a) - oracle.llh_to_ecef(*b)))


def test_oracle_three_stations_2d_recovers_the_transmitter():
    rd = range_diffs(STATION_LLH, TX)
    out, rms, status, it = oracle.solve_ls(STATION_LLH, rd, init_llh=[41.26, -96.02, 400.0], dims=2)
    assert status == 0 and rms < 1e-6 and it < 60
    assert dist_m(out, TX) < 1e-3          # north_star: position within 1 m
    # ... where the reference's own solver, from its own start point, is kilometres off after 10 half steps
    ref, st, _ = oracle.solve_tdoa(STATION_LLH, rd)
    assert st == 0 and dist_m(ref, TX) > 100.0


def test_oracle_sixteen_stations_3d():
    st = ring(16)
    rd = range_diffs(st, TX)
    out, rms, status, it = oracle.solve_ls(st, rd, init_llh=None, dims=3)
    assert status == 0 and rms < 1e-6
    assert dist_m(out, TX) < 1e-2
    noisy = rd + np.random.default_rng(1).normal(0, 50e-9 * C_LIGHT, rd.size)   # 50 ns timing noise
    out2, rms2, status2, _ = oracle.solve_ls(st, noisy, init_llh=[41.26, -96.02, TX[2]], dims=2)
    assert status2 == 0 and 5.0 < rms2 < 30.0
    assert dist_m([out2[0], out2[1], TX[2]], TX) < 60.0


@pytest.mark.gpu
def test_gpu_least_squares_matches_oracle():
    import tdoa_b200 as T
    rng = np.random.default_rng(7)
    with T.Engine(T.MODE_BINARY) as e:
        # 3 stations, 2-D, a batch of transmitters seeded from the grid arg-min
        txs = np.array([[41.20 + 0.05 * rng.uniform(-1, 1), -96.00 + 0.05 * rng.uniform(-1, 1), 400.0] for _ in range(33)])
        rds = np.array([range_diffs(STATION_LLH, t) for t in txs])
        init = np.tile([41.26, -96.02, 400.0], (len(txs), 1))
        out, rms, status, iters = e.solve_ls(STATION_LLH, rds, init_llh=init, dims=2)
        for k in range(len(txs)):
            want, wrms, wst, wit = oracle.solve_ls(STATION_LLH, rds[k], init_llh=init[k], dims=2)
            assert status[k] == wst == 0
            assert dist_m(out[k], want) < 1e-5 and abs(rms[k] - wrms) < 1e-6
            assert dist_m(out[k], txs[k]) < 1e-3
        # 16 stations, 3-D, noisy, default start (station mean)
        st = ring(16)
        rd = range_diffs(st, TX) + rng.normal(0, 50e-9 * C_LIGHT, 120)
        got, grms, gst, git = e.solve_ls(st, rd, dims=3)
        want, wrms, wst, wit = oracle.solve_ls(st, rd, dims=3)
        assert gst == wst == 0 and dist_m(got, want) < 1e-4 and abs(grms - wrms) < 1e-6
        # chained behind the dense grid: the arg-min cell is the start point
        desc = [41.26 - 0.25, -96.02 - 0.25, 0.0005, 0.0005, 1000, 1000, 400.0]
        cell, cost, idx = e.grid(st, desc, rd)
        fine, frms, fst, _ = e.solve_ls(st, rd, init_llh=cell, dims=2)
        assert fst == 0 and frms <= np.sqrt(cost / 120) + 1e-9 and dist_m(fine, cell) < 60.0
        with pytest.raises(T.TdoaError):
            e.solve_ls(STATION_LLH, rds[0], dims=3)   # elevation from 3 stations
