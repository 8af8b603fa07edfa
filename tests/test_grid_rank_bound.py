"""The ranked grid arg-min (solve.cu: k_grid_rank ... k_grid_pick) rests on one inequality: the expanded cost

    f = [S sum q^2 - (sum q)^2] - 2 sum_k q_k c_k + sum rd^2,   q_k = r_k - r_0,  c_k = sum_{i<k} rd_ik - sum_{j>k} rd_kj

differs from the statement's cost (sum over pairs of ((r_j - r_i) - rd_ij)^2, added pair by pair) by less than

    tol = 4e-13 (2 S sum q^2 + 4 max|q| sum|rd| + sum rd^2).

Then the statement's arg-min survives the cut f - tol <= min (f + tol) and the statement settles the survivors.
This file restates both forms in numpy -- f64 operation by operation as the kernels do them (the FMA chains in
exact rational arithmetic rounded once, via Python fractions on a sample; plain f64 on the bulk) -- and checks the
inequality over geometries far more hostile than config 5: ranges from metres to 10^6 m, range differences from
exact to garbage.  CPU only; the GPU comparison with the exhaustive kernel is tests/test_gpu_parity.py.
"""
from fractions import Fraction

import numpy as np

KAPPA = 4e-13


def table_row(rd, S):
    """grid_rank_row (solve.cu, host): c_k, sum rd^2, 2 sum |rd|."""
    c = np.zeros(S)
    b = 0.0
    p = 0
    for i in range(S):
        for j in range(i + 1, S):
            c[j] += rd[p]
            c[i] -= rd[p]
            b += rd[p] * rd[p]
            p += 1
    return c, b, 2.0 * float(np.sum(np.abs(rd)))


def statement(r, rd):
    cost = 0.0
    p = 0
    S = len(r)
    for i in range(S):
        for j in range(i + 1, S):
            e = (r[j] - r[i]) - rd[p]
            cost = cost + e * e
            p += 1
    return cost


def fma(a, b, c):
    """One correctly rounded a * b + c (what __fma_rn does), through exact rationals."""
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def ranked(r, c, b, c1, exact_fma):
    S = len(r)
    q = [rk - r[0] for rk in r]
    sq = 0.0
    sl = 0.0
    qm = 0.0
    for x in q:
        sq = fma(x, x, sq) if exact_fma else x * x + sq
        sl = sl + x
        qm = max(qm, abs(x))
    a = S * sq - sl * sl
    t0 = 2.0 * S * sq
    mid = 0.0
    for x, ck in zip(q, c):
        mid = fma(x, ck, mid) if exact_fma else x * ck + mid
    f = fma(-2.0, mid, a + b) if exact_fma else -2.0 * mid + (a + b)
    tol = KAPPA * (t0 + ((fma(2.0 * qm, c1, b)) if exact_fma else (2.0 * qm * c1 + b)))
    return f, tol


def cases(rng, n):
    for _ in range(n):
        S = int(rng.integers(2, 17))
        scale = 10.0 ** rng.uniform(0, 6)                       # ranges from metres to 1000 km
        r = scale * rng.uniform(0.05, 1.0, S)
        true_rd = np.array([r[j] - r[i] for i in range(S) for j in range(i + 1, S)])
        kind = rng.integers(0, 4)
        if kind == 0:
            rd = true_rd.copy()                                  # the cell IS the transmitter: cost ~ 0 from 1e12-size terms
        elif kind == 1:
            rd = true_rd + rng.normal(0, 15.0, true_rd.size)     # timing noise
        elif kind == 2:
            rd = true_rd * rng.uniform(0.5, 1.5)                 # a far-away cell
        else:
            rd = scale * rng.uniform(-3, 3, true_rd.size)        # garbage
        yield r, rd, S


def test_rounding_bound_holds_f64():
    rng = np.random.default_rng(20261019)
    worst = 0.0
    for r, rd, S in cases(rng, 20000):
        c, b, c1 = table_row(rd, S)
        f, tol = ranked(list(r), c, b, c1, exact_fma=False)
        err = abs(f - statement(r, rd))
        assert err <= tol, (err, tol, S)
        worst = max(worst, err / tol)
    assert worst < 0.1      # the constant has an order of magnitude to spare


def test_rounding_bound_holds_with_fused_chains():
    """The kernels' FMA chains (one rounding per step) instead of numpy's multiply-then-add."""
    rng = np.random.default_rng(7)
    for r, rd, S in cases(rng, 600):
        c, b, c1 = table_row(rd, S)
        f, tol = ranked([float(x) for x in r], [float(x) for x in c], b, c1, exact_fma=True)
        assert abs(f - statement(r, rd)) <= tol


def test_cut_keeps_the_statements_argmin():
    """A small grid end to end in numpy: every cell whose f - tol is under min (f + tol) is kept, the statement
    picks among the kept ones, and that is the statement's arg-min over all cells -- duplicates (ties) included."""
    rng = np.random.default_rng(3)
    for trial in range(40):
        S = int(rng.integers(3, 9))
        st = rng.uniform(-3e4, 3e4, (S, 3))
        cells = rng.uniform(-4e4, 4e4, (400, 3))
        cells[200:] = cells[:200]                      # every cell twice: exact ties, the lower index has to win
        tx = cells[int(rng.integers(0, 200))] + rng.normal(0, 30.0, 3)
        rt = np.linalg.norm(st - tx, axis=1)
        rd = np.array([rt[j] - rt[i] for i in range(S) for j in range(i + 1, S)]) + rng.normal(0, [0.0, 20.0][trial & 1], S * (S - 1) // 2)
        c, b, c1 = table_row(rd, S)
        rs = [np.linalg.norm(st - x, axis=1) for x in cells]
        costs = np.array([statement(r, rd) for r in rs])
        ft = [ranked(list(r), c, b, c1, exact_fma=False) for r in rs]
        bound = min(f + t for f, t in ft)
        kept = [k for k, (f, t) in enumerate(ft) if f - t <= bound]
        best = min(kept, key=lambda k: (costs[k], k))
        want = int(np.flatnonzero(costs == costs.min())[0])
        assert best == want
        assert len(kept) <= 8                          # the cut is sharp: the minimum, its twin, perhaps a neighbour
