#!/usr/bin/env python3
"""developer tool: pinned host -> device copy bandwidth of this box — what bounds the end-to-end numbers.
  python tools/h2d_peak.py            one GPU, several copy sizes
  python tools/h2d_peak.py --concurrent 1,2,4,8
                                       N processes, one per GPU, copying at the same time (the e2e leg of bench.py under
                                       torchrun does exactly this): per-GPU and aggregate GB/s per N, one JSON line each"""
import argparse
import json
import multiprocessing as mp
import os
import time


def worker(gpu, n_procs, mb, reps, start_at, q):
    os.environ["CUDA_VISIBLE_DEVICES"] = str(gpu)
    import torch
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    while time.time() < start_at:      # all processes start their timed copies together
        pass
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    q.put((gpu, n * reps / (e0.elapsed_time(e1) / 1e3) / 1e9))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--concurrent", default="")
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=16)
    a = ap.parse_args()
    if not a.concurrent:
        import torch
        for mb in (48, 256, 1024, 3072):
            n = mb << 20
            h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            h.fill_(1)
            d = torch.empty(n, dtype=torch.uint8, device="cuda")
            for _ in range(2):
                d.copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            reps = max(2, 4096 // mb)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                d.copy_(h, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            print(f"H2D {mb} MiB x{reps}: {n * reps / (e0.elapsed_time(e1) / 1e3) / 1e9:.1f} GB/s")
        return
    ctx = mp.get_context("spawn")
    for n_procs in [int(x) for x in a.concurrent.split(",")]:
        q = ctx.Queue()
        start_at = time.time() + 12.0      # CUDA init + pinning 1 GiB per process takes a few seconds
        ps = [ctx.Process(target=worker, args=(g, n_procs, a.mb, a.reps, start_at, q)) for g in range(n_procs)]
        for p in ps:
            p.start()
        res = sorted(q.get(timeout=300) for _ in ps)
        for p in ps:
            p.join()
        per = [round(r[1], 1) for r in res]
        print(json.dumps({"concurrent_gpus": n_procs, "copy_mib": a.mb, "reps": a.reps, "per_gpu_gbs": per,
                          "aggregate_gbs": round(sum(per), 1), "min_gbs": min(per)}), flush=True)


if __name__ == "__main__":
    main()
