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

import csv
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
            raise RuntimeError(f"failed to load stations: failed to open CSV file: {exc}") from exc
        with f:
            records = list(csv.reader(f))
        for i, rec in enumerate(records[1:]):  # header skipped (:66)
            if len(rec) != 4:
                raise RuntimeError(f"failed to load stations: invalid CSV format at line {i + 2}")
            try:
                lat, lon, elev = float(rec[1]), float(rec[2]), float(rec[3])
            except ValueError as exc:
                raise RuntimeError(f"failed to load stations: invalid number at line {i + 2}: {exc}") from exc
            st = Station(rec[0], lat, lon, elev)
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
        try:
            raw = np.fromfile(filename, dtype=np.uint8)
        except OSError as exc:
            raise RuntimeError(f"failed to open file: {exc}") from exc
        n = raw.size // 2
        self._print(f"File size: {raw.size} bytes, samples: {n}")
        self.engine(n_stations).load_u8(slot, raw)
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

    # -- processor.go:739-929
    def process_tdoa(self, dat_files: List[str]):
        if len(dat_files) < 3:
            raise RuntimeError(f"need at least 3 collector stations, got {len(dat_files)}")
        P = self._print
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
                self.load_iq_data(slot, fn, S)
            except RuntimeError as exc:
                raise RuntimeError(f"failed to load data from {fn}: {exc}") from exc
            stations.append(st)
            P("Loaded collector: %s at %.6f°, %.6f°, %.1fm" % (st.name, st.latitude, st.longitude, st.elevation))
        llh = np.array([s.llh for s in stations], np.float64)
        P("\nBaseline distances (3D):")
        base = eng.baselines(llh)
        pairs = [(i, j) for i in range(S) for j in range(i + 1, S)]
        for (i, j), d in zip(pairs, base):
            P("%s - %s: %.2f km" % (stations[i].name, stations[j].name, d / 1000))
        fs = eng.cfg.sample_rate
        results = {}
        for kind, label in ((N.KIND_REF, "REF"), (N.KIND_TGT, "TGT")):
            if kind == N.KIND_REF:
                P("\n=== REFERENCE SIGNAL CORRELATION TEST ===")
            else:
                P("\n=== TARGET SIGNAL CORRELATION TEST ===")
            peaks = eng.xcorr(kind)[0]
            tds = []
            for (i, j), pk in zip(pairs, peaks):
                td = float(pk["lag"]) / fs
                tds.append(td)
                P("%s %s - %s: delay=%d samples (%.3f μs), correlation=%.6f"
                  % (label, stations[i].name, stations[j].name, pk["lag"], td * 1e6, pk["corr"]))
            results[label] = (peaks, tds)
        ref_td, tgt_td = results["REF"][1], results["TGT"][1]
        if self.mode == N.MODE_SOURCE:
            tds = tgt_td  # processor.go:853: target differences only
        else:
            # shipped binary: "REFERENCE SIGNAL SYNCHRONIZATION" corrected = tgt - ref
            tds = [t - r_ for t, r_ in zip(tgt_td, ref_td)]
        rds = [td * SPEED_OF_LIGHT for td in tds]  # :899-903
        P("\n=== TDOA GEOLOCATION ===")
        P("Time differences (μs): " + "".join("%.3f " % (td * 1e6) for td in tds))
        P("Range differences (m): " + "".join("%.1f " % rd for rd in rds))
        try:
            lat, lon, elev = self.solve_tdoa(stations, rds)
        except RuntimeError as exc:
            raise RuntimeError(f"TDOA solution failed: {exc}") from exc
        P("\n*** CALCULATED TRANSMITTER LOCATION ***")
        P("Latitude:  %.6f°" % lat)
        P("Longitude: %.6f°" % lon)
        P("Elevation: %.1f m" % elev)
        return {"ref": results["REF"][0], "tgt": results["TGT"][0], "time_differences": tds,
                "range_differences": rds, "position": (lat, lon, elev)}


def main(argv=None) -> int:
    """processor <ref_freq_hz> <target_freq_hz> <stations.csv> <dat...>  (processor.go:1047-1076)"""
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 6:
        print("Usage: processor <reference_freq_hz> <target_freq_hz> <stations.csv> <collector1.dat> "
              "<collector2.dat> <collector3.dat> [...]")
        return 1
    try:
        p = TDOAProcessor(float(argv[0]), float(argv[1]), argv[2])
        p.process_tdoa(argv[3:])
    except RuntimeError as exc:
        print(f"TDOA processing failed: {exc}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
