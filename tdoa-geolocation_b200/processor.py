"""Host-side mirror of the reference's processing-stage interface (processor.go):
same names, argument meaning, pair order, error behaviour and stdout records, with
every numeric step delegated to the CUDA engine through the C ABI.

    p = TDOAProcessor(162400000, 92300000, "lat-lon-table.csv")
    p.process_tdoa(["sim-kx0u-1.dat", "sim-n3pay-1.dat", "sim-kf0mtl-1.dat"])

Reference seams (file:line): NewTDOAProcessor :36, loadStations :52,
getStationFromFilename :110, loadIQData :166, extract* :208/:241, crossCorrelate :619,
ProcessTDOA :739, solveTDOA :932.
"""
from __future__ import annotations

import math
import os
import sys
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N

SPEED_OF_LIGHT = 299792458.0  # processor.go:899


@dataclass
class Station:  # processor.go:15-20
    name: str
    latitude: float
    longitude: float
    elevation: float

    @property
    def llh(self):
        return (self.latitude, self.longitude, self.elevation)


class TDOAProcessor:
    """processor.go:29-33 TDOAProcessor, backed by one tdoa_engine."""

    def __init__(self, ref_freq: float, target_freq: float, csv_path: str, mode: int = N.MODE_BINARY,
                 out=None, **engine_overrides):
        self.reference_freq = float(ref_freq)
        self.target_freq = float(target_freq)
        self.stations = {}
        self.ref_station: Optional[Station] = None
        self.mode = mode
        self.out = out if out is not None else sys.stdout
        self._overrides = engine_overrides
        self._engine: Optional[N.Engine] = None
        self.load_stations(csv_path)

    def _print(self, *a, **k):
        print(*a, file=self.out, **k)

    # -- processor.go:52-107
    def load_stations(self, csv_path: str) -> None:
        try:
            f = open(csv_path, newline="")
        except OSError as exc:
            raise RuntimeError("failed to load stations: failed to open CSV file: open %s: %s"
                               % (csv_path, _go_strerror(exc))) from exc
        with f:
            try:
                records = _read_csv_all(f.read())
            except ValueError as exc:
                raise RuntimeError(f"failed to load stations: failed to read CSV: {exc}") from exc
        for i, rec in enumerate(records[1:]):  # header skipped (:66)
            if len(rec) != 4:
                raise RuntimeError(f"failed to load stations: invalid CSV format at line {i + 2}")
            vals = []
            for what, cell in zip(("latitude", "longitude", "elevation"), rec[1:]):
                try:
                    vals.append(_parse_float(cell))
                except ValueError as exc:
                    raise RuntimeError(f"failed to load stations: invalid {what} at line {i + 2}: {exc}") from exc
            st = Station(rec[0], *vals)
            self.stations[rec[0]] = st
            if rec[0] == "%.0f" % self.reference_freq:  # :96
                self.ref_station = st
        if self.ref_station is None:
            raise RuntimeError("failed to load stations: reference frequency %.0f not found in stations"
                               % self.reference_freq)
        self._print("Loaded %d stations including reference %.0f MHz" % (len(self.stations),
                                                                         self.reference_freq / 1e6))

    # -- processor.go:110-122.  The reference ranges over a Go map (random order); ties
    # between nested names are broken here by preferring the longest name.
    def get_station_from_filename(self, filename: str) -> Station:
        base = os.path.basename(filename)
        for name in sorted(self.stations, key=lambda s: (-len(s), s)):
            if name in base:
                return self.stations[name]
        raise RuntimeError(f"could not identify station from filename: {filename}")

    # -- engine management
    def engine(self, n_stations: int) -> N.Engine:
        if self._engine is None or self._engine.n_stations != n_stations:
            if self._engine is not None:
                self._engine.close()
            self._engine = N.Engine(self.mode, n_stations=n_stations, **self._overrides)
        return self._engine

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    # -- processor.go:166-205 (bytes go to the GPU; the complex64 samples stay there)
    def load_iq_data(self, slot: int, filename: str, n_stations: int) -> int:
        self._print(f"Loading I/Q data from: {filename}")
        # the reference prints the size between opening and reading (:176-183); a file that cannot
        # be opened is reported by tdoa_load_file with the reference's text
        have = os.path.exists(filename) and os.access(filename, os.R_OK)
        if have:
            size = os.stat(filename).st_size
            self._print(f"File size: {size} bytes, samples: {size // 2}")
        try:
            # tdoa_load_file: streamed to the device through pinned staging, never held on the host
            n = self.engine(n_stations).load_file(slot, filename)
        except N.TdoaError as exc:
            raise RuntimeError(exc.message) from exc   # "failed to open file: ..." as processor.go:170-172
        if not have:
            self._print(f"File size: {2 * n} bytes, samples: {n}")
        self._print(f"Successfully loaded {n} complex samples")
        return n

    # -- processor.go:619-643
    def cross_correlate(self, signal1, signal2) -> Tuple[int, float]:
        pk = self.engine(max(2, self._engine.n_stations if self._engine else 3)).cross_correlate(signal1, signal2)
        return pk.lag, pk.corr

    # -- processor.go:932-1020
    def solve_tdoa(self, stations: Sequence[Station], range_differences) -> Tuple[float, float, float]:
        if len(stations) < 3 or len(range_differences) < 2:
            raise RuntimeError("need at least 3 stations for TDOA")
        llh = np.array([s.llh for s in stations], np.float64)
        out, status, _ = self.engine(len(stations)).solve(llh, np.asarray(range_differences, np.float64))
        if status != 0:
            raise RuntimeError("singular Jacobian matrix")  # :997-999
        return float(out[0]), float(out[1]), float(out[2])

    # -- stdout of the shipped binary while it works on one pair (ELF 0x49cd40, 0x49d6a0)
    _BRANCH_TEXT = ("Strong FM signal - using instantaneous frequency correlation approach",
                    "Moderate signal - envelope correlation approach",
                    "Weak signal - standard processing with timing preservation")

    def _print_pair_binary(self, sig1, sig2, pk, first_corr, fs, max_lag, block, sanity):
        P = self._print
        P("=== Cross-Correlation Analysis ===")
        if sig1["n"] == 0 or sig2["n"] == 0:   # processor.go:622-625
            P("Warning: Empty signals for correlation")
            return
        P("\n--- Signal Preprocessing ---")
        for k, sg in ((1, sig1), (2, sig2)):
            P("Preprocessing Signal %d signal (%d samples)" % (k, sg["n"]))
            P("Initial signal power: %.9f" % sg["power0"])
            P(self._BRANCH_TEXT[sg["branch"]])
            P("Removed DC bias: %.6f + %.6fi" % (sg["dc_re"], sg["dc_im"]))
            if sg["branch"] == 2:
                P("Bandpass filter: %.1f - %.1f Hz (at %.0f Hz sample rate)" % (100.0, 200000.0, 2000000.0))
            if sg["power1"] > 0:   # processor.go:338-340, :349
                P("Normalized signal power: %.6f → 1.000000" % sg["power1"])
        P("\n--- Time Domain Correlation ---")
        P("Performing time domain correlation")
        tl, sl = min(sig1["n"], sig2["n"]), max(sig1["n"], sig2["n"])
        P("Template: %d samples, Signal: %d samples" % (tl, sl))
        if tl == sl:
            P("Reduced template to %d samples to allow %d sample delay search" % (tl - max_lag, max_lag))
        P("Using coherent integration with %d-sample blocks" % block)
        n_lags = max_lag if tl == sl else max(1, min(max_lag, sl - tl))
        P("Time domain progress: 0/%d (coherent blocks: %d)" % (n_lags, pk["n_blocks"]))
        first_lag = int(pk["first_lag"])
        P("Time domain correlation: %.6f at delay %d samples" % (first_corr, first_lag))
        if sanity > 0 and first_lag > sanity:
            P("WARNING: Delay %d samples (%.1f μs) exceeds reasonable range for baseline distances"
              % (first_lag, first_lag / fs * 1e6))
            P("Maximum expected delay: 56.7 μs for 17 km baseline")
            P("This suggests correlation algorithm found wrong peak")
            if int(pk["flags"]) & 1:
                P("Found better peak within reasonable range: delay=%d samples (%.1f μs), correlation=%.6f"
                  % (pk["lag"], pk["lag"] / fs * 1e6, pk["corr"]))
        P("\n--- Result: Time Domain with Preprocessing ---")
        P("Correlation: %.6f at delay %d samples" % (pk["corr"], pk["lag"]))

    def _print_pair_source(self, sig1, sig2, pk, max_lag, block):
        """stdout of processor.go while it works on one pair (crossCorrelate :619-643, preprocessSignal
        :469-499, enhanceWeakSignal :437-466, timeDomainCorrelation :646-736)."""
        P = self._print
        P("=== Cross-Correlation Analysis ===")
        if sig1["n"] == 0 or sig2["n"] == 0:   # :622-625
            P("Warning: Empty signals for correlation")
            return
        P("\n--- Signal Preprocessing ---")
        bp = "Bandpass filter: %.1f - %.1f Hz (at %.0f Hz sample rate)"
        for k, sg in ((1, sig1), (2, sig2)):
            P("Preprocessing Signal %d signal (%d samples)" % (k, sg["n"]))
            P("Initial signal power: %.9f" % sg["power0"])
            if sg["branch"] == 2:   # power < 0.001 (:476)
                P("Detected very weak signal - applying aggressive filtering")
                P("Enhancing weak signal: Signal %d" % k)
                P("Removed DC bias: %.6f + %.6fi" % (sg["dc_re"], sg["dc_im"]))
                for lo, hi in ((57.5, 62.5), (117.5, 122.5), (975000.0, 1000000.0), (100.0, 40000.0)):   # :446-448, :457
                    P(bp % (lo, hi, 2000000.0))
            else:
                P("Standard signal processing")
                P("Removed DC bias: %.6f + %.6fi" % (sg["dc_re"], sg["dc_im"]))
                P(bp % (500.0, 50000.0, 2000000.0))   # :489
            if sg["power1"] > 0:   # :338-340, :349
                P("Normalized signal power: %.6f → 1.000000" % sg["power1"])
        P("\n--- Time Domain Correlation ---")
        P("Performing time domain correlation")
        tl, sl = min(sig1["n"], sig2["n"]), max(sig1["n"], sig2["n"])
        P("Template: %d samples, Signal: %d samples" % (tl, sl))
        n_lags = max(1, min(max_lag, sl - tl))   # :668-675
        P("Using coherent integration with %d-sample blocks" % block)
        # :729-731: every 2000th delay, carriage return instead of a new line
        progress = "".join("Time domain progress: %d/%d (coherent blocks: %d)\r" % (d, n_lags, pk["n_blocks"])
                           for d in range(0, n_lags, 2000))
        P(progress + "\nTime domain correlation: %.6f at delay %d samples" % (pk["corr"], pk["lag"]))
        P("\n--- Result: Time Domain with Preprocessing ---")
        P("Correlation: %.6f at delay %d samples" % (pk["corr"], pk["lag"]))

    # -- processor.go:739-929; in MODE_BINARY every stdout line of the shipped binary
    def process_tdoa(self, dat_files: List[str]):
        if len(dat_files) < 3:
            raise RuntimeError(f"need at least 3 collector stations, got {len(dat_files)}")
        P = self._print
        binary = self.mode == N.MODE_BINARY
        P("Processing TDOA for target frequency %.3f MHz" % (self.target_freq / 1e6))
        r = self.ref_station
        P("Reference: %s at %.6f°, %.6f°, %.1fm" % (r.name, r.latitude, r.longitude, r.elevation))
        S = len(dat_files)
        eng = self.engine(S)
        stations = []
        for slot, fn in enumerate(dat_files):
            try:
                st = self.get_station_from_filename(fn)
            except RuntimeError as exc:
                raise RuntimeError(f"failed to identify station for {fn}: {exc}") from exc
            try:
                n = self.load_iq_data(slot, fn, S)
            except RuntimeError as exc:
                raise RuntimeError(f"failed to load data from {fn}: {exc}") from exc
            b = n // 3  # processor.go:211-236, :244-265
            n_ref, n_tgt = (n, n) if n < 3 else (2 * b, b)   # fewer than 3 samples: returned unchanged
            P("Extracting reference signal from dual-frequency data")
            if n < 3:
                P("Warning: Data too small for dual-frequency extraction")
            else:
                P("Total samples: %d, block size: %d" % (n, b))
                P("Extracted %d reference samples from blocks 1 and 3" % (2 * b))
            P("Extracting target signal from dual-frequency data")
            if n < 3:
                P("Warning: Data too small for dual-frequency extraction")
            else:
                P("Total samples: %d, block size: %d" % (n, b))
                P("Extracted %d target samples from block 2" % b)
            # processor.go:772-783 (testChunkSize 2 000 000; the shipped binary: 1 000 000) = the engine's chunk_samples
            chunk = eng.cfg.chunk_samples
            if n_ref > chunk:
                P("Using test chunk: %d samples (%.1f ms)" % (chunk, chunk / 2e6 * 1000))
            if n_tgt > chunk:
                P("Using target test chunk: %d samples (%.1f ms)" % (chunk, chunk / 2e6 * 1000))
            P("Coherent integration time: %.0f ms (expecting ~%.1f dB processing gain)"
              % (chunk / 2e6 * 1000, 10 * math.log10(chunk / 100000)))
            stations.append(st)
            P("Loaded collector: %s at %.6f°, %.6f°, %.1fm" % (st.name, st.latitude, st.longitude, st.elevation))
        llh = np.array([s.llh for s in stations], np.float64)
        P("\nBaseline distances (3D):")
        base = eng.baselines(llh)
        pairs = [(i, j) for i in range(S) for j in range(i + 1, S)]
        for (i, j), d in zip(pairs, base):
            P("%s - %s: %.2f km" % (stations[i].name, stations[j].name, d / 1000))
        fs = eng.cfg.sample_rate
        # the whole numeric path in one engine call (tdoa_process): both pair loops, the time and
        # range differences and solveTDOA are queued on the GPU without a host round trip; what
        # follows only prints what the reference prints, in its order
        done = eng.process(llh)
        results = {}
        for kind, label in ((N.KIND_REF, "REF"), (N.KIND_TGT, "TGT")):
            if kind == N.KIND_REF:
                P("\n=== REFERENCE SIGNAL CORRELATION TEST ===")
                # processor.go:813 prints the frequency as a literal; the shipped binary formats it
                P("Testing weak %.1f MHz NOAA weather signal:" % (self.reference_freq / 1e6) if binary
                  else "Testing weak 162.4 MHz NOAA weather signal:")
            else:
                P("\n=== TARGET SIGNAL CORRELATION TEST ===")
                P("Testing strong %.1f MHz FM broadcast signal:" % (self.target_freq / 1e6) if binary
                  else "Testing strong 92.3 MHz FM broadcast signal:")   # processor.go:833
            peaks = done["ref" if kind == N.KIND_REF else "tgt"]
            info, first = eng.xcorr_info(kind)
            tds = []
            for p_idx, ((i, j), pk) in enumerate(zip(pairs, peaks)):
                td = float(pk["lag"]) / fs
                tds.append(td)
                if binary:
                    self._print_pair_binary(info[i], info[j], pk, float(first[p_idx]), fs, eng.cfg.max_lag,
                                            eng.cfg.block_size, eng.cfg.sanity_lag)
                else:
                    self._print_pair_source(info[i], info[j], pk, eng.cfg.max_lag, eng.cfg.block_size)
                P("%s %s - %s: delay=%d samples (%.3f μs), correlation=%.6f"
                  % (label, stations[i].name, stations[j].name, pk["lag"], td * 1e6, pk["corr"]))
            results[label] = (peaks, tds)
        ref_td, tgt_td = results["REF"][1], results["TGT"][1]
        if not binary:
            tds = tgt_td  # processor.go:853: target differences only
            P("\n=== CORRELATION COMPARISON ===")   # :855-858
            P("Reference signal (162.4 MHz): Generally weaker correlation")
            P("Target signal (92.3 MHz): Should show stronger correlation")
            P("Using target signal for TDOA calculation")
        else:
            # shipped binary: corrected = target - reference, pair by pair
            P("\n=== REFERENCE SIGNAL SYNCHRONIZATION ===")
            P("Using reference signal to synchronize collector timing...")
            for k, t in enumerate(ref_td):
                P("Reference timing offset %d: %.3f μs" % (k, t * 1e6))
            P("\n=== APPLYING TIMING CORRECTIONS TO TARGET SIGNAL ===")
            tds = [t - r_ for t, r_ in zip(tgt_td, ref_td)]
            for k, (t, r_, c) in enumerate(zip(tgt_td, ref_td, tds)):
                P("Target delay %d: %.3f μs (raw) - %.3f μs (ref offset) = %.3f μs (corrected)"
                  % (k, t * 1e6, r_ * 1e6, c * 1e6))
            P("\n=== CORRELATION COMPARISON ===")
            P("Reference signal (%.1f MHz): Used for timing synchronization" % (self.reference_freq / 1e6))
            P("Target signal (%.1f MHz): Corrected with reference timing offsets" % (self.target_freq / 1e6))
            P("Using corrected target signal for TDOA calculation")
        rds = [td * SPEED_OF_LIGHT for td in tds]  # :899-903
        P("\nTDOA triangulation:")
        # both revisions print the first three here, whatever the number of pairs (:869-879)
        P(("Corrected time differences: " if binary else "Time differences: ") + ", ".join("%.3f μs" % (td * 1e6) for td in tds[:3]))
        P(("Corrected distance differences: " if binary else "Distance differences: ") + ", ".join("%.1f m" % rd for rd in rds[:3]))
        P("\nDiagnostic test with example delays:")  # processor.go:882-889
        P("Simulating 10 μs, 5 μs, -3 μs delays...")
        for k, us in enumerate((10.0, 5.0, -3.0)):
            P("Test delay %d: %.1f μs → %.1f m" % (k + 1, us, us * 1e-6 * SPEED_OF_LIGHT))
        P("\n=== TDOA GEOLOCATION ===")
        P("Time differences (μs): " + "".join("%.3f " % (td * 1e6) for td in tds))
        P("Range differences (m): " + "".join("%.1f " % rd for rd in rds))
        if binary:
            # shipped binary: measurements beyond 1.2 x 17 km are dropped before its solver
            # (which then aborts on every input, SURVEY.md finding 4; the fix below is
            # processor.go's solveTDOA on the unfiltered differences)
            P("Validating range differences against baseline distances...")
            limit, valid = 20400.0, 0
            for k, rd in enumerate(rds):
                if abs(rd) <= limit:
                    P("VALID: Range difference %d: %.1fm (within ±%.1fm limit)" % (k, rd, limit))
                    valid += 1
                else:
                    P("FILTERING OUT: Range difference %d: %.1fm exceeds expected maximum %.1fm" % (k, rd, limit))
                    P("This measurement is unreliable and will be excluded")
            # the binary's own solveTDOA (ELF 0x4a0360) from here on: tdoa_solve_binary
            if valid < 2:
                raise RuntimeError("TDOA solution failed: insufficient valid measurements: only %d of %d range differences are reliable"
                                   % (valid, len(rds)))
            P("Using %d of %d range difference measurements" % (valid, len(rds)))
            e0, e1, e2 = (ecef(*s.llh) for s in stations[:3])
            area = 0.5 * abs((e1[0] - e0[0]) * (e2[1] - e0[1]) - (e2[0] - e0[0]) * (e1[1] - e0[1]))
            P("Station geometry triangle area: %.1f m²" % area)
            if area < 1e7:
                P("WARNING: Poor station geometry (small triangle area)")
                P("This may cause TDOA solution instability")
            m = llh[:3].mean(axis=0)
            P("Initial guess: %.6f°, %.6f°, %.1fm" % (m[0], m[1], m[2]))
        # same arithmetic on the device: dt = delay / fs, (target - reference), * c
        if self.mode in (N.MODE_SOURCE, N.MODE_BINARY):
            assert np.array_equal(np.asarray(rds, np.float64), done["range_differences"])
        if binary:
            pos, b_status, _, n_iter, conv, trace = eng.solve_binary(llh, np.asarray(rds, np.float64))
            if b_status == 2:
                raise RuntimeError("TDOA solution failed: no valid range difference measurements remain")
            for k, t in enumerate(trace):
                P("Iteration %d: det=%.2e, residuals=[%.1f, %.1f]" % (k, t[0], t[1], t[2]))
                code, last = int(t[4]), b_status == 3 and k == n_iter - 1
                if code == 1:
                    P("Large step detected (%.1fm) - limiting to %.1fm" % (t[3], 1000.0 * (1000.0 / t[3] * 0.7)))
                if code >= 2 or last:
                    P("Singular matrix detected (det=%.2e) - trying alternative approach" % t[0])
                if code >= 2:
                    P("Using single equation approach (equation %d)" % (code - 1))
                if last:
                    raise RuntimeError("TDOA solution failed: singular Jacobian matrix at iteration %d (det=%.2e)" % (k, t[0]))
                if k == 9:
                    P("Maximum iterations reached")
            if conv:
                P("Converged after %d iterations" % n_iter)
            done = dict(done, position=pos, status=0)
        else:
            # processor.go:957, :971, :998, :1013 (solveTDOA ran inside tdoa_process; fix_iters = the
            # iteration it stopped at: converged, singular, or 10 = all ten steps taken)
            m = llh[:3].sum(axis=0) / 3.0
            P("Initial guess: %.6f°, %.6f°, %.1fm" % (m[0], m[1], m[2]))
            if done["status"] != 0:
                raise RuntimeError("TDOA solution failed: singular Jacobian matrix at iteration %d" % done["iters"])
            P("Converged after %d iterations" % done["iters"] if done["iters"] < 10 else "Maximum iterations reached")
        lat, lon, elev = (float(x) for x in done["position"])
        P("\n*** CALCULATED TRANSMITTER LOCATION ***")
        P("Latitude:  %.6f°" % lat)
        P("Longitude: %.6f°" % lon)
        P("Elevation: %.1f m" % elev)
        return {"ref": results["REF"][0], "tgt": results["TGT"][0], "time_differences": tds,
                "range_differences": rds, "position": (lat, lon, elev)}


def ecef(lat, lon, h):
    """processor.go:125-148 latLonToECEF (WGS-84), host copy used only for the printed triangle area."""
    import math
    a, f = 6378137.0, 1.0 / 298.257223563
    e2 = 2 * f - f * f
    la, lo = math.radians(lat), math.radians(lon)
    n = a / math.sqrt(1 - e2 * math.sin(la) ** 2)
    return ((n + h) * math.cos(la) * math.cos(lo), (n + h) * math.cos(la) * math.sin(lo),
            (n * (1 - e2) + h) * math.sin(la))


def _go_strerror(exc: OSError) -> str:
    """Go's text for an errno: the C library's with a lower-case first letter."""
    msg = os.strerror(exc.errno) if exc.errno else str(exc)
    return msg[:1].lower() + msg[1:]


def _parse_float(text: str) -> float:
    """strconv.ParseFloat(s, 64) as far as the station table and the command line need it."""
    bad = ValueError('strconv.ParseFloat: parsing "%s": invalid syntax' % text)
    if not text or text != text.strip() or "_" in text:
        raise bad
    try:
        v = float(text)
    except ValueError:
        try:
            v = float.fromhex(text) if text.lower().lstrip("+-").startswith("0x") else None
        except ValueError:
            v = None
        if v is None:
            raise bad from None
    if v in (float("inf"), float("-inf")) and "inf" not in text.lower():
        raise ValueError('strconv.ParseFloat: parsing "%s": value out of range' % text)
    return v


def _read_csv_all(text: str):
    """encoding/csv ReadAll as far as a station table needs it (the C++ mirror's readCsvAll):
    quoted fields with "" for a quote, blank lines skipped, every record as long as the first."""
    records = []
    for line_no, line in enumerate(text.split("\n"), 1):
        if line.endswith("\r"):
            line = line[:-1]
        if not line:
            continue
        rec, i = [], 0
        while True:
            cell = ""
            if i < len(line) and line[i] == '"':
                i += 1
                while True:
                    if i >= len(line):
                        raise ValueError('record on line %d; parse error on line %d, column %d: extraneous or missing " in quoted-field'
                                         % (line_no, line_no, i + 1))
                    if line[i] == '"':
                        if i + 1 < len(line) and line[i + 1] == '"':
                            cell += '"'
                            i += 2
                            continue
                        i += 1
                        break
                    cell += line[i]
                    i += 1
                if i < len(line) and line[i] != ",":
                    raise ValueError('parse error on line %d, column %d: extraneous or missing " in quoted-field' % (line_no, i + 1))
            else:
                while i < len(line) and line[i] != ",":
                    if line[i] == '"':
                        raise ValueError('parse error on line %d, column %d: bare " in non-quoted-field' % (line_no, i + 1))
                    cell += line[i]
                    i += 1
            rec.append(cell)
            if i >= len(line):
                break
            i += 1
            if i == len(line):
                rec.append("")
                break
        if records and len(rec) != len(records[0]):
            raise ValueError("record on line %d: wrong number of fields" % line_no)
        records.append(rec)
    return records


def _fatalf(msg: str) -> int:
    """log.Fatalf: '2006/01/02 15:04:05 ' + message on stderr, exit status 1."""
    import time
    sys.stdout.flush()
    print(time.strftime("%Y/%m/%d %H:%M:%S"), msg, file=sys.stderr)
    return 1


def main(argv=None, prog: str = "processor") -> int:
    """processor <ref_freq_hz> <target_freq_hz> <csv_file> <dat_file1> ...  (processor.go:1047-1076)"""
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 4:   # :1048-1052
        print(f"Usage: {prog} <ref_freq_hz> <target_freq_hz> <csv_file> <dat_file1> [dat_file2] [dat_file3] ...")
        print("Example: ./processor 162400000 101700000 lat-lon-table.csv kx0u-data.dat n3pay-data.dat kf0mtl-data.dat")
        return 1
    try:
        ref_freq = _parse_float(argv[0])
    except ValueError as exc:
        return _fatalf(f"Invalid reference frequency: {exc}")
    try:
        tgt_freq = _parse_float(argv[1])
    except ValueError as exc:
        return _fatalf(f"Invalid target frequency: {exc}")
    try:
        p = TDOAProcessor(ref_freq, tgt_freq, argv[2])
    except RuntimeError as exc:
        return _fatalf(f"Failed to create processor: {exc}")
    try:
        p.process_tdoa(argv[3:])
    except RuntimeError as exc:
        return _fatalf(f"TDOA processing failed: {exc}")
    finally:
        p.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
