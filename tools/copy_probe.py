"""How much host<->device copy bandwidth do N ranks get when they copy at the same time?

Evidence for the e2e figure of bench.py (173 MB in + 173 MB out per rank and step through pinned host memory):
every rank runs plain cudaMemcpyAsync H2D and D2H of that size concurrently on two streams, all ranks
start together (barrier), and rank 0 prints the per-rank and aggregate GB/s.  No kernel of this repo runs.

    python -m torch.distributed.run --nproc-per-node N tools/copy_probe.py   (or plain python for N = 1)
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 172748800
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for it in range(2):                                   # first pass = warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        gbs = 20 * nbytes / dt / 1e9
        t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
        if world > 1:
            lo, hi, sm = t.clone(), t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(sm)
            res[mode] = {"per_rank_min_GBs": float(lo), "per_rank_max_GBs": float(hi), "aggregate_GBs_per_direction": float(sm)}
        else:
            res[mode] = {"per_rank_min_GBs": gbs, "per_rank_max_GBs": gbs, "aggregate_GBs_per_direction": gbs}
    if rank == 0:
        print(json.dumps({"ranks": world, "bytes_per_copy": nbytes, "copies": res,
                          "note": "'both' = H2D and D2H in flight together, GB/s per direction"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
