#!/usr/bin/env python
"""BASELINE configs 3 and 4 with the windows sharded across the GPUs of one box
(tdoa-geolocation_b200/sharding.py): rank r takes windows w = r (mod world), runs every pair
of its windows locally -- no data-path collective -- and one NCCL all_gather of the 32-byte
peak records leaves every rank with every window.  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port P tools/sharded_windows.py [--config 3|4] [--reps K]

Every rank synthesises the same captures (same seeds) on its own GPU; timing = CUDA events
around the sharded sweep + gather, max over ranks.  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tools"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import configs_bench as cb  # noqa: E402
import tdoa_b200 as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--block", type=int, default=66_666_666)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = 2_000_000
    if args.config == 4:
        n_st, L = 16, 2000
        stations = cb.ring_stations(16)
    else:
        n_st, L = 3, 50_000
        stations = bench.STATION_LLH
    delays, _ = cb.delays_for(stations)
    caps = cb.synth(dev, n_st, args.block, delays)
    want = [int(delays[j] - delays[i]) for i in range(n_st) for j in range(i + 1, n_st)]
    nw = args.block // W
    with T.Engine(T.MODE_EXTENDED, n_stations=n_st, max_lag=L, fast_demod=1, device=local) as e:
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        for k in range(n_st):
            e.load_u8_device(k, caps[k].data_ptr(), caps[k].numel(), keep=caps[k])

        def sweep():
            mine, peaks = T.sharding.local_xcorr(e, T.KIND_TGT, 0, W, nw, W, rank, world)
            return T.sharding.gather_peaks(mine, peaks, nw, e.n_pairs, device=dev if world > 1 else None)

        full = sweep()  # warm-up + the result that is checked
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            full = sweep()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev, dtype=torch.float64)
        digest = torch.tensor([float(np.int64(full["lag"]).sum()), float(full["corr"].sum())], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            lo, hi = digest.clone(), digest.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(lo, hi))
        else:
            same = True
    ok = all([int(x) for x in full[w]["lag"]] == want for w in range(nw))
    if rank == 0:
        t = float(ms.item()) / 1e3
        print(json.dumps({"config": args.config, "n_gpus": world, "windows": nw, "pairs": len(want), "ms_per_sweep": t * 1e3,
                          "pair_msamples_per_s": nw * len(want) * W / t / 1e6, "fixes_per_s": nw / t,
                          "every_window_recovers_the_injected_delays": bool(ok), "all_ranks_hold_the_same_records": same,
                          "collective": "one all_gather of %d-byte records per sweep" % (32 * len(want) * ((nw + world - 1) // world))}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
