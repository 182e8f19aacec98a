#!/usr/bin/env python
"""What the BOX can feed: N ranks (one per GPU) each copy 120 MB of pinned host memory to their GPU, all at once,
plain cudaMemcpyAsync, nothing else running.  Prints the aggregate GB/s (max-over-ranks time) -- the ceiling that
bench.py's end-to-end number is compared with (`h2d_ceiling_gbs`, `e2e_frac_of_h2d_ceiling`; bench.py measures the
same thing in-line with this function).

    python tools/h2d_ceiling.py                                            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/h2d_ceiling.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    import coverage_b200 as cov
    c = bench.Ctx()
    c.torch, c.dist = torch, dist
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(c.local)
    note = bench.bind_to_gpu_numa_node(c.local) if c.world > 1 else "single rank: not bound"
    if c.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", c.local))
    eng = cov.CoverageEngine(c.local)
    c.stream = torch.cuda.Stream(device=c.local)
    eng.set_stream(c.stream.cuda_stream)
    out = {}
    for mb, sets in ((120, 1), (120, 4), (120, 8)):
        out[f"{mb}MB_x{sets}_buffers"] = bench.h2d_ceiling(c, eng, nbytes=mb * 1_000_000, reps=16, sets=sets)
    # ... and with the path's result traffic (17 B per candidate = 17 MB per 120 MB copy) flowing back at the same time
    out["120MB_x4_buffers_with_17MB_d2h"] = bench.h2d_ceiling(c, eng, nbytes=120_000_000, reps=16, sets=4, d2h_bytes=17_000_000)
    if c.rank == 0:
        print(json.dumps({"n_gpus": c.world, "aggregate_h2d_gbs": out, "per_gpu_gbs": {k: v / c.world for k, v in out.items()},
                          "host_placement": note}))
    eng.close()
    if c.world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
