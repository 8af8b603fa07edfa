"""Host-side multi-GPU logic (tdoa-geolocation_b200/sharding.py) on CPU: window/pair
partitioning and the single gather of peak records, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import tdoa_b200 as T
from importlib import import_module

N = import_module("tdoa-geolocation_b200._native")
S = T.sharding


def test_window_partition_covers_everything_once():
    for n in (0, 1, 7, 33, 66):
        for world in (1, 2, 3, 8):
            seen = sorted(w for r in range(world) for w in S.shard_windows(n, r, world))
            assert seen == list(range(n))
            sizes = [len(S.shard_windows(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        S.shard_windows(4, 2, 2)


def test_rotating_cursor_balances_back_to_back_calls():
    """The engine's rule (tdoa_shard_windows, host arithmetic of libtdoa_b200.so): window w belongs to
    rank (cursor + w) % world and the cursor advances by the windows of every sharded call, so that
    66 REF windows followed by 33 TGT windows over 8 ranks (BASELINE config 4) leave every rank
    with 12 or 13 windows instead of 9 + 5 on rank 0."""
    world, cursor, load = 8, 0, [0] * 8
    for n in (66, 33):
        seen = []
        for r in range(world):
            ws = S.shard_windows(n, r, world, cursor)
            assert all((cursor + w) % world == r for w in ws)
            load[r] += len(ws)
            seen += ws
        assert sorted(seen) == list(range(n))
        cursor = (cursor + n) % world
    assert sorted(set(load)) == [12, 13] and sum(load) == 99
    first, count = N.shard_windows_rule(5, 3, 4, 2)   # rank 3 of 4, window 0 belongs to rank 2
    assert (first, count) == (1, 1)
    assert N.shard_windows_rule(2, 1, 8, 6) == (3, 0)  # nothing for this rank
    with pytest.raises(ValueError):
        N.shard_windows_rule(4, 2, 2, 0)


def test_window_runs_reconstruct_the_windows():
    for ws in ([], [5], [0, 2, 4, 6], [1, 9, 17], [0, 1, 2, 10, 20, 30, 31]):
        out = []
        for first, count, stride in S.window_runs(ws):
            out += [first + k * stride for k in range(count)]
        assert out == list(ws)
    assert S.window_runs([3, 11, 19, 27]) == [(3, 4, 8)]


def test_pair_partition():
    assert sorted(p for r in range(8) for p in S.shard_pairs(120, r, 8)) == list(range(120))


class FakeEngine:
    """Stands in for the CUDA engine: records the calls, returns recognisable peaks."""
    n_pairs = 3

    def xcorr(self, kind, win_start, win_len, n_windows, hop):
        out = np.zeros((n_windows, 3), N.PEAK_DTYPE)
        for w in range(n_windows):
            start = win_start + w * hop
            out[w]["lag"] = [start, start + 1, start + 2]
            out[w]["corr"] = start / 1000.0
        return out


def _worker(rank, world, port, n_windows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hop, start = 1000, 50
    mine, peaks = S.local_xcorr(FakeEngine(), T.KIND_TGT, start, 800, n_windows, hop, rank, world)
    full = S.gather_peaks(mine, peaks, n_windows, 3)
    q.put((rank, mine, full["lag"].tolist(), full["corr"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_windows", [5, 8])
def test_gather_world_size_2_gloo(n_windows):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_windows, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_lag = [[50 + 1000 * w + k for k in range(3)] for w in range(n_windows)]
    for rank, mine, lag, corr in res:
        assert mine == list(range(rank, n_windows, 2))
        assert lag == want_lag                      # every rank holds every window, in order
        assert np.allclose(np.array(corr)[:, 0], [(50 + 1000 * w) / 1000.0 for w in range(n_windows)])
