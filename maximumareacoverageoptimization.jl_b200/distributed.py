"""Candidate sharding over the GPUs of one box, one process per GPU (torch.distributed).

Candidates are independent, so the path shards with NO data-path collective: rank r
evaluates the contiguous slice [r*ceil(B/G), (r+1)*ceil(B/G)) on its own GPU with the grid and the
parameter block replicated.  The only exchange is at the end and only if the caller wants it:
  * `gather_objectives`  all-gather of the per-rank objective slices (8*B/G bytes each), or
  * `argmin_pair`        a 16-byte (min objective, global index) pair per rank, all-gathered and
                         reduced on the host (ties -> smallest index), for a MADS poll winner.
Works with any torch.distributed backend (NCCL over NVLink on the GPU box; gloo in the CPU tests,
where the evaluation callable is a stand-in because the CUDA library needs a device).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist


def shard_range(B: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous slice of rank `rank`: (first, count) with ceil(B/world) per rank (the last ranks may
    get fewer or none) -- the same rule as cov_multi_eval_batch in the C ABI."""
    per = (B + world - 1) // world
    b0 = min(B, per * rank)
    return b0, min(B, b0 + per) - b0


def _device_for_backend():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def eval_sharded(evaluate, X: np.ndarray, gather: bool = True):
    """`evaluate(X_slice) -> obj_slice` (e.g. `lambda S: engine.eval_batch(S)["obj"]`) on this rank's
    slice of the globally known candidate matrix X.  Returns the full objective vector on every rank
    when gather=True, else (first, obj_slice)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    B = X.shape[0]
    b0, bn = shard_range(B, world, rank)
    mine = np.asarray(evaluate(X[b0:b0 + bn]), dtype=np.float64) if bn else np.zeros(0)
    if not gather:
        return b0, mine
    return gather_objectives(mine, B)


def gather_objectives(obj_slice: np.ndarray, B: int) -> np.ndarray:
    world = dist.get_world_size()
    per = (B + world - 1) // world
    dev = _device_for_backend()
    buf = torch.full((per,), math.nan, dtype=torch.float64, device=dev)
    buf[:len(obj_slice)] = torch.from_numpy(np.ascontiguousarray(obj_slice)).to(dev)
    out = torch.empty((world * per,), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf) if dev.type == "cuda" else dist.all_gather(
        list(out.view(world, per).unbind(0)), buf)
    return out.cpu().numpy()[:B] if per * world == B else np.concatenate(
        [out.view(world, per)[r, :shard_range(B, world, r)[1]].cpu().numpy() for r in range(world)])


def argmin_pair(obj_slice: np.ndarray, first: int, feasible: np.ndarray | None = None):
    """Global (min objective, index) from per-rank slices; infeasible candidates count as +inf
    (extreme barrier).  Returns (inf, -1) when nothing is feasible anywhere."""
    world = dist.get_world_size()
    v = np.asarray(obj_slice, dtype=np.float64)
    if feasible is not None:
        v = np.where(np.asarray(feasible, dtype=bool), v, math.inf)
    v = np.where(np.isnan(v), math.inf, v)
    if len(v) and np.isfinite(v).any():
        k = int(np.argmin(v))
        pair = (float(v[k]), float(first + k))
    else:
        pair = (math.inf, -1.0)
    dev = _device_for_backend()
    mine = torch.tensor(pair, dtype=torch.float64, device=dev)
    allp = [torch.empty(2, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(allp, mine)
    best, idx = math.inf, -1
    for t in allp:
        o, i = t.cpu().tolist()
        if i >= 0 and (idx < 0 or o < best or (o == best and i < idx)):
            best, idx = o, int(i)
    return best, idx


def exchange_scratch():
    """Reusable buffers for exchange_winner on the current backend's device: pinned host staging on both sides of
    the all-gather (two asynchronous 16-byte / 16*world-byte copies and ONE synchronise per exchange)."""
    world = dist.get_world_size()
    dev = _device_for_backend()
    pin = dev.type == "cuda"
    return {"pair_h": torch.empty(2, dtype=torch.float64, pin_memory=pin),
            "pair": torch.empty(2, dtype=torch.float64, device=dev),
            "gathered": torch.empty(2 * world, dtype=torch.float64, device=dev),
            "gathered_h": torch.empty(2 * world, dtype=torch.float64, pin_memory=pin)}


def exchange_winner(best_obj: float, best_idx: int, first: int, scratch=None):
    """The end-of-poll exchange when every rank already holds its own winner (cov_eval_batch_best / cov_argmin reduce
    it on the device): all-gather of one (objective, GLOBAL index) pair per rank -- 16 bytes each -- and the same
    tie rule as argmin_pair (smallest objective, then smallest index).  best_idx < 0: this rank has no feasible
    candidate.  scratch: exchange_scratch() to reuse between calls.
    Returns (objective, global index, winning rank); (inf, -1, -1) when no rank has a feasible candidate."""
    world = dist.get_world_size()
    sc = scratch if scratch is not None else exchange_scratch()
    sc["pair_h"][0] = best_obj if best_idx >= 0 else math.inf
    sc["pair_h"][1] = float(first + best_idx) if best_idx >= 0 else -1.0
    sc["pair"].copy_(sc["pair_h"], non_blocking=True)
    if sc["pair"].device.type == "cuda":
        dist.all_gather_into_tensor(sc["gathered"], sc["pair"])
        sc["gathered_h"].copy_(sc["gathered"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    else:
        dist.all_gather(list(sc["gathered"].view(world, 2).unbind(0)), sc["pair"])
        sc["gathered_h"].copy_(sc["gathered"])
    rows = sc["gathered_h"].view(world, 2).tolist()
    best, idx, who = math.inf, -1, -1
    for r, (o, i) in enumerate(rows):
        if i >= 0 and (idx < 0 or o < best or (o == best and i < idx)):
            best, idx, who = o, int(i), r
    return best, idx, who
