"""The analyzer restatement in the oracle (orc_analyze_samples: fast_analyzer.go /
analyzer.go, O(M^2) DFT as the reference) against an independent numpy statement of the
same definitions (np.fft for the spectrum).  The reference ships no analyzer binary and no
fixtures for this path, so this cross-check is what pins the oracle ("parity unpinned" by
the reference itself)."""
import numpy as np
import pytest

from oracle import oracle
from helpers import fm_capture


def numpy_quality(sig: np.ndarray, fast: bool) -> dict:
    i, q = sig[0::2].astype(np.float64), sig[1::2].astype(np.float64)
    n = i.size
    out = {"total_samples": n, "i_avg": i.sum() / n, "q_avg": q.sum() / n}
    out["i_std"] = np.sqrt((i * i).sum() / n - out["i_avg"] ** 2)
    out["q_std"] = np.sqrt((q * q).sum() / n - out["q_avg"] ** 2)
    mag = np.hypot(out["i_std"], out["q_std"])
    out["power_db"] = -100.0 if (fast and mag <= 1e-10) else 20 * np.log10(mag)
    out["has_clipping"] = int(i.min() == 0 or i.max() == 255 or q.min() == 0 or q.max() == 255)
    out["has_overload"] = int(out["i_std"] < 2 or out["q_std"] < 2)
    m = min(8192 if fast else 16384, n)
    s0 = (n - m) // 2
    ii, qq = i[s0:s0 + m], q[s0:s0 + m]
    k = np.arange(m) / (m - 1)
    if fast:
        x = ((ii - 127.5) + 1j * (qq - 127.5)) / 127.5 * (0.5 - 0.5 * np.cos(2 * np.pi * k))
    else:
        w = 0.35875 - 0.48829 * np.cos(2 * np.pi * k) + 0.14128 * np.cos(4 * np.pi * k) - 0.01168 * np.cos(6 * np.pi * k)
        x = ((ii - ii.mean()) + 1j * (qq - qq.mean())) / 127.5 * w
    psd = np.abs(np.fft.fft(x)) ** 2
    srt = np.sort(psd)
    sig_thr = srt[int(0.9 * m)]
    sp = psd[psd >= sig_thr].mean()
    if fast:
        noise = psd[(psd < sig_thr) & (psd <= srt[int(0.4 * m)])]
        npw = noise.mean() if noise.size else 0.0
    else:
        npw = srt[:int(0.5 * m)].mean()
    out["snr_db"] = 10 * np.log10(sp / npw) if (npw > 0 and sp > npw) else -20.0
    if not fast:
        out["dc_offset"] = np.hypot(out["i_avg"] - 127.5, out["q_avg"] - 127.5)
        out["iq_imbalance"] = abs(out["i_std"] - out["q_std"]) / max(out["i_std"], out["q_std"])
        z = np.concatenate([[1], (sig != 0).astype(np.int8)])  # closed runs only: a run needs a terminating non-zero
        idx = np.flatnonzero(z)
        longest = int(np.max(np.diff(idx) - 1)) if idx.size > 1 else 0
        out["has_dead_zones"] = int(longest > 1000)
        out["has_noise"] = int(out["i_std"] > 60 or out["q_std"] > 60)
    return out


@pytest.mark.parametrize("fast", [True, False])
def test_oracle_analyzer_against_numpy(fast):
    rng = np.random.default_rng(11)
    raw = fm_capture(12000, (0, 7, 3), (0, 7, 3), seed=5)[0]
    raw[30000:31200] = 0          # a dead zone inside block 2 (bytes 24000..48000 -> samples of TGT)
    raw[100] = 255
    sigs = [raw[:24000], raw[24000:48000], rng.integers(0, 256, 6000, dtype=np.uint8)]
    for s in sigs:
        got = oracle.analyze_samples(s, fast)
        want = numpy_quality(s, fast)
        for k, w in want.items():
            if isinstance(w, (int, np.integer)):
                assert got[k] == w, (k, got[k], w)
            else:
                assert got[k] == pytest.approx(w, rel=1e-9, abs=1e-7), (k, got[k], w)
