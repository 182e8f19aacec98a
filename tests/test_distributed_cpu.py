"""CPU-only, world_size 2, gloo: the host logic of candidate sharding (the N > 1 path)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coverage_b200 import distributed as D
    rng = np.random.default_rng(0)
    X = rng.random((B, 6))
    f = lambda S: np.sum((S - 0.3) ** 2, axis=1)  # noqa: E731  stand-in for engine.eval_batch(S)["obj"]
    full = D.eval_sharded(f, X, gather=True)
    b0, mine = D.eval_sharded(f, X, gather=False)
    feas = (np.arange(b0, b0 + len(mine)) % 3) != 0
    best = D.argmin_pair(mine, b0, feas)
    none = D.argmin_pair(mine, b0, np.zeros(len(mine), dtype=bool))
    # every rank already holds its winner (what cov_eval_batch_best returns): the 16-byte exchange
    v = np.where(feas, mine, np.inf)
    k = int(np.argmin(v)) if len(v) and np.isfinite(v).any() else -1
    win = D.exchange_winner(float(v[k]) if k >= 0 else np.inf, k, b0)
    nobody = D.exchange_winner(np.inf, -1, b0)
    q.put((rank, full, b0, len(mine), best, none, win, nobody))
    dist.destroy_process_group()


def test_shard_range_rule(cov):
    from coverage_b200.distributed import shard_range
    for B in (0, 1, 7, 8, 9, 1000003):
        for G in (1, 2, 4, 8):
            parts = [shard_range(B, G, r) for r in range(G)]
            assert sum(n for _, n in parts) == B
            pos = 0
            for b0, n in parts:
                assert b0 == min(pos, B) and n >= 0
                pos += n
            assert max(n for _, n in parts) == (B + G - 1) // G or B == 0


def test_gloo_world2_shard_gather_argmin():
    B, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    X = rng.random((B, 6))
    want = np.sum((X - 0.3) ** 2, axis=1)
    masked = np.where(np.arange(B) % 3 != 0, want, np.inf)
    for rank, full, b0, n, best, none, win, nobody in res:
        assert np.array_equal(full, want)
        assert (b0, n) == ((0, 501) if rank == 0 else (501, 500))
        assert best == (masked.min(), int(np.argmin(masked)))
        assert none == (np.inf, -1)
        g = int(np.argmin(masked))
        assert win == (masked.min(), g, 0 if g < 501 else 1) and nobody == (np.inf, -1, -1)
