"""Multi-GPU sharding of the processing stage: one process per GPU, units
are independent (window, pair) correlations and per-window fixes, so the compute path
needs no exchange.  Windows are dealt round-robin to ranks (rank r takes windows
w = r (mod world)); every rank copies only what its windows need, preprocesses each
station-window once and runs all pairs locally.  The only collective is one gather of
the 32-byte peak records (KB-scale, latency-bound) over NCCL/NVLink -- or gloo in the
CPU tests of this host logic.

The reference has no counterpart (single process, processor.go:816-850 is a plain double
loop); pair order inside a window stays i<j lexicographic = argv order.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from ._native import PEAK_DTYPE, comm_unique_id, shard_windows_rule


def shard_windows(n_windows: int, rank: int, world: int, cursor: int = 0) -> List[int]:
    """Window indices owned by `rank`: window w belongs to rank (cursor + w) % world (round-robin:
    balances a trailing partial group).  The rule is the engine's own (tdoa_shard_windows), the
    one its in-library sharded tdoa_xcorr applies."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    first, count = shard_windows_rule(n_windows, rank, world, cursor)
    return list(range(first, first + count * world, world))


def init_engine_comm(engine, group=None) -> None:
    """One process per GPU under torch.distributed: rank 0 makes the NCCL id (tdoa_comm_unique_id),
    the process group carries its 128 bytes to the others, every rank joins (tdoa_comm_init).  From
    then on engine.xcorr over >= 2 windows is sharded and gathered inside the library."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.tensor(list(comm_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0, group=group)
    engine.comm_init(bytes(t.cpu().tolist()), rank, world)


def window_runs(windows: Sequence[int]) -> List[Tuple[int, int, int]]:
    """Group a rank's windows into (first, count, stride) runs so each run is ONE
    tdoa_xcorr call (win_start = start + first*hop, n_windows = count, hop = stride*hop)."""
    runs: List[Tuple[int, int, int]] = []
    ws = list(windows)
    i = 0
    while i < len(ws):
        if i + 1 == len(ws):
            runs.append((ws[i], 1, 1))
            break
        stride = ws[i + 1] - ws[i]
        j = i + 1
        while j + 1 < len(ws) and ws[j + 1] - ws[j] == stride:
            j += 1
        runs.append((ws[i], j - i + 1, stride))
        i = j + 1
    return runs


def shard_pairs(n_pairs: int, rank: int, world: int) -> List[int]:
    """Pair sharding for the single-window configs (needs every station on every rank;
    pays only when there are more pairs than GPUs, e.g. 120 pairs of 16 stations)."""
    return list(range(rank, n_pairs, world))


def local_xcorr(engine, kind: int, win_start: int, win_len: int, n_windows: int, hop: int, rank: int,
                world: int) -> Tuple[List[int], np.ndarray]:
    """Run this rank's share of an n_windows sweep; returns (window indices, peaks[len][P])."""
    mine = shard_windows(n_windows, rank, world)
    out = np.zeros((len(mine), engine.n_pairs), PEAK_DTYPE)
    row = 0
    for first, count, stride in window_runs(mine):
        res = engine.xcorr(kind, win_start + first * hop, win_len, count, stride * hop if count > 1 else 0)
        out[row:row + count] = res
        row += count
    return mine, out


def gather_peaks(local_windows: Sequence[int], local_peaks: np.ndarray, n_windows: int, n_pairs: int,
                 group=None, device=None) -> np.ndarray:
    """The single collective of the path: all ranks end up with peaks[n_windows][P] in
    window order.  Records travel as raw bytes (the C ABI's 32-byte tdoa_peak)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    full = np.zeros((n_windows, n_pairs), PEAK_DTYPE)
    if world == 1:
        full[list(local_windows)] = local_peaks
        return full
    per_rank = (n_windows + world - 1) // world
    buf = np.zeros((per_rank, n_pairs), PEAK_DTYPE)
    buf[:len(local_windows)] = local_peaks
    t = torch.from_numpy(buf.view(np.uint8).reshape(-1).copy())
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    for r, o in enumerate(outs):
        rec = o.cpu().numpy().view(PEAK_DTYPE).reshape(per_rank, n_pairs)
        ws = shard_windows(n_windows, r, world)
        full[ws] = rec[:len(ws)]
    return full
