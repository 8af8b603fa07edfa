"""More than one GPU behind the C ABI (engine_multi.cu): windows dealt over the
ranks of a communicator inside tdoa_xcorr, the records meeting in one ncclAllGather.  A one-rank
communicator runs the same code on a one-GPU box; the two-rank cases need two GPUs (gpurun --gpus 2)
and skip otherwise."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import tdoa_b200 as T
from helpers import fm_capture

pytestmark = pytest.mark.gpu

W, HOP, NW, L = 30000, 25000, 7, 300


def captures():
    return fm_capture(220000, (0, 3, 8), (0, 20, 41), seed=17)


def load_all(e, raws):
    for k, r in enumerate(raws):
        e.load_u8(k, r)


def reference_table(raws, kind, mode=T.MODE_EXTENDED):
    with T.Engine(mode, max_lag=L) as e:
        load_all(e, raws)
        return e.xcorr(kind, 1000, W, NW, HOP)


def same_records(a, b):
    for name in ("lag", "corr", "frac", "first_lag", "n_blocks"):
        assert np.array_equal(a[name], b[name]), name


def test_xcorr_device_leaves_the_records_on_the_gpu():
    """tdoa_xcorr_device (the entry point an NCCL send buffer is filled through): same records as
    tdoa_xcorr, left in caller-owned device memory."""
    raws = captures()
    want = reference_table(raws, T.KIND_TGT)
    buf = torch.zeros(NW * 3 * 32, dtype=torch.uint8, device="cuda")
    with T.Engine(T.MODE_EXTENDED, max_lag=L) as e:
        load_all(e, raws)
        e.xcorr_device(T.KIND_TGT, buf.data_ptr(), 1000, W, NW, HOP)
        e.synchronize()
        got = buf.cpu().numpy().view(T.sharding.PEAK_DTYPE).reshape(NW, 3)
    same_records(got, want)


def test_one_rank_communicator_takes_the_sharded_path():
    """tdoa_comm_init with world = 1: gather buffer, ncclAllGather, window-order kernel -- and a
    cursor that advances call by call -- on one GPU.  Records equal the plain call's."""
    raws = captures()
    want_t, want_r = reference_table(raws, T.KIND_TGT), reference_table(raws, T.KIND_REF)
    with T.Engine(T.MODE_EXTENDED, max_lag=L) as e:
        load_all(e, raws)
        e.comm_init(T.comm_unique_id(), 0, 1)
        assert e.comm_rank() == (0, 1)
        same_records(e.xcorr(T.KIND_TGT, 1000, W, NW, HOP), want_t)
        same_records(e.xcorr(T.KIND_REF, 1000, W, NW, HOP), want_r)
        buf = torch.zeros(NW * 3 * 32, dtype=torch.uint8, device="cuda")
        e.xcorr_device(T.KIND_TGT, buf.data_ptr(), 1000, W, NW, HOP)
        e.synchronize()
        same_records(buf.cpu().numpy().view(T.sharding.PEAK_DTYPE).reshape(NW, 3), want_t)
        with pytest.raises(T.TdoaError):
            e.comm_init(T.comm_unique_id(), 0, 1)   # one communicator per engine


def test_xcorr_windows_is_the_two_pair_loops():
    """tdoa_xcorr_windows: both pair loops over windows in one call.  On one GPU it is tdoa_xcorr(REF) followed
    by tdoa_xcorr(TGT); behind a communicator (one rank here) the windows of both kinds are dealt as one list and
    gathered at the end -- the same records either way, unequal window counts and a zero count included."""
    raws = captures()
    want_r, want_t = reference_table(raws, T.KIND_REF), reference_table(raws, T.KIND_TGT)
    with T.Engine(T.MODE_EXTENDED, max_lag=L) as e:
        load_all(e, raws)
        r, t = e.xcorr_windows(1000, W, NW, NW - 2, HOP)
        same_records(r, want_r)
        same_records(t, want_t[:NW - 2])
        e.comm_init(T.comm_unique_id(), 0, 1)
        r, t = e.xcorr_windows(1000, W, NW, NW - 2, HOP)
        same_records(r, want_r)
        same_records(t, want_t[:NW - 2])
        r, t = e.xcorr_windows(1000, W, 0, NW, HOP)
        assert r.shape[0] == 0
        same_records(t, want_t)


def test_n_devices_beyond_the_box_is_refused():
    n = torch.cuda.device_count()
    with pytest.raises(T.TdoaError, match="n_devices"):
        T.Engine(T.MODE_BINARY, n_devices=n + 1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_process_two_devices():
    """tdoa_config.n_devices = 2: one process, the engine forwards the loads to its peer and deals
    the windows of tdoa_xcorr over both GPUs (one host thread per device)."""
    raws = captures()
    with T.Engine(T.MODE_EXTENDED, max_lag=L, n_devices=2) as e:
        load_all(e, raws)
        for kind in (T.KIND_TGT, T.KIND_REF, T.KIND_TGT):      # the cursor rotates: 7 windows, 2 ranks
            same_records(e.xcorr(kind, 1000, W, NW, HOP), reference_table(raws, kind))
        one = e.xcorr(T.KIND_TGT, 1000, W, 1, 0)               # a single window runs on device 0
        same_records(one, reference_table(raws, T.KIND_TGT)[:1])
        r, t = e.xcorr_windows(1000, W, NW, NW - 2, HOP)       # 7 + 5 windows dealt as one list of 12
        same_records(r, reference_table(raws, T.KIND_REF))
        same_records(t, reference_table(raws, T.KIND_TGT)[:NW - 2])


def _rank_main(rank, world, id_path, q):
    torch.cuda.set_device(rank)
    raws = captures()
    with T.Engine(T.MODE_EXTENDED, max_lag=L, device=rank) as e:
        if rank == 0:
            uid = T.comm_unique_id()
            with open(id_path + ".tmp", "wb") as f:
                f.write(uid)
            os.replace(id_path + ".tmp", id_path)
        else:
            import time
            while not os.path.exists(id_path):
                time.sleep(0.05)
            uid = open(id_path, "rb").read()
        load_all(e, raws)
        e.comm_init(uid, rank, world)
        out = [e.xcorr(kind, 1000, W, NW, HOP) for kind in (T.KIND_TGT, T.KIND_REF)]
        r2, t2 = e.xcorr_windows(1000, W, NW, NW, HOP)
        q.put((rank, [o.tobytes() for o in out + [t2, r2]]))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_process_per_gpu_two_ranks():
    """tdoa_comm_unique_id / tdoa_comm_init: two processes, one GPU each; tdoa_xcorr is collective and
    both ranks end with the whole table, equal to a one-GPU run's."""
    raws = captures()
    want = [reference_table(raws, T.KIND_TGT), reference_table(raws, T.KIND_REF)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as td:
        procs = [ctx.Process(target=_rank_main, args=(r, 2, os.path.join(td, "id"), q)) for r in range(2)]
        for p in procs:
            p.start()
        res = dict(q.get(timeout=300) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    for r in range(2):
        for got, w in zip(res[r], want + want):   # tdoa_xcorr per kind, then tdoa_xcorr_windows: the same tables
            same_records(np.frombuffer(got, T.sharding.PEAK_DTYPE).reshape(NW, 3), w)
