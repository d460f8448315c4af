"""Replica sharding across the GPUs of one box.

Replicas are independent (no edge crosses a replica: reference `build_edge_idx`, GNN_Models.py:926-942), so each
rank owns a contiguous block of ceil(R / world) replicas, evaluates / minimises / integrates it with no
inter-GPU traffic, and the only collective of the path is one gather of the per-replica energies at the end."""
from __future__ import annotations

from typing import List, Tuple

import torch


def shard_bounds(n_rep: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the replicas owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(n_rep), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def interleaved_solvents(n_rep: int, n_solvents: int) -> List[int]:
    """Solvent id of replica r = r mod n_solvents (Simulation/helper_functions.py:738-743), so that every
    contiguous shard sees all solvent embeddings."""
    return [r % n_solvents for r in range(n_rep)]


def gather_energies(e_local: torch.Tensor, n_rep: int, world: int, rank: int) -> torch.Tensor:
    """all_gather of the per-replica energies (ragged shards padded to the largest one). Returns E[n_rep] on the
    device of `e_local` (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    if world == 1:
        return e_local
    width = max(shard_bounds(n_rep, world, r)[1] - shard_bounds(n_rep, world, r)[0] for r in range(world))
    buf = torch.zeros(width, dtype=e_local.dtype, device=e_local.device)
    buf[: e_local.numel()] = e_local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_rep, world, r)
        out.append(parts[r][: hi - lo])
    return torch.cat(out)
