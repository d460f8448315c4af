"""Replica sharding across ranks: contiguous blocks, no data-path collective, one gather of
energies at the end.  Host logic only -- runs on CPU with the gloo backend, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnnimplicitsolvent_b200 import sharding


def test_shard_bounds_cover_all_replicas_once():
    for R in (1, 7, 64, 4096, 19968):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(R, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == R
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_solvent_interleaving_per_shard():
    """C3: solvent id = replica mod 39 (helper_functions.py:738-743); every shard of 2496 replicas sees all 39."""
    sids = sharding.interleaved_solvents(19968, 39)
    for rank in range(8):
        lo, hi = sharding.shard_bounds(19968, 8, rank)
        assert hi - lo == 2496
        assert set(sids[lo:hi]) == set(range(39))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, R, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(R, world, rank)
    # stand-in for the per-rank engine: the "energy" of replica r is a known function of r
    e_local = torch.arange(lo, hi, dtype=torch.float32) * 0.5 - 3.0
    e_all = sharding.gather_energies(e_local, R, world, rank)
    if rank == 0:
        np.save(out, e_all.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("R", [5, 64, 1001])
def test_gather_energies_gloo_world2(tmp_path, R):
    out = str(tmp_path / "e.npy")
    mp.spawn(_worker, args=(2, _free_port(), R, out), nprocs=2, join=True)
    np.testing.assert_array_equal(np.load(out), np.arange(R, dtype=np.float32) * 0.5 - 3.0)
