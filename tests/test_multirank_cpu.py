"""CPU tests of the N>1 plumbing with the gloo backend (world_size 2): the rank partition, the broadcast of
the communicator id bytes, and the Comm::Sum semantics (partial value/derivatives/virial of every rank add
up to the single-rank result) that both the NCCL combine in the library and the plugin's comm.Sum rely on."""
import os
import socket

import numpy as np
import pytest

import plumed2_b200 as P
from helpers import water_box
from oracle import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the communicator id travels as an opaque 128-byte blob from rank 0 (bench.py does the same over NCCL)
        blob = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(blob, src=0)
        assert blob[0] == bytes(range(128))
        pos, box = water_box(n, 100.0, seed=77)
        pbc = O.make_pbc(box)
        sw = O.make_switch("RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8")
        nl = O.NeighborList(O.NL_SINGLELIST, n, 0, cutoff=1.0, stride=10)
        nl.update(pbc, pos, fast=True)
        v, d, w, _ = O.coordination(nl, pbc, True, sw, pos, rank=rank, nranks=world)
        buf = torch.from_numpy(np.concatenate([d.ravel(), w.ravel(), [v]]))
        dist.all_reduce(buf)  # Communicator::Sum of 3N + 9 + 1 doubles (CoordinationBase.cpp:218-224)
        lo, cnt = P.shard_range(n, rank, world)
        sl = torch.zeros(world, 2, dtype=torch.int64)
        sl[rank, 0], sl[rank, 1] = lo, cnt
        dist.all_reduce(sl)
        if rank == 0:
            np.save(os.path.join(out_dir, "sum.npy"), buf.numpy())
            np.save(os.path.join(out_dir, "slices.npy"), sl.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_sum_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    n, world = 1500, 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "sum.npy")
    pos, box = water_box(n, 100.0, seed=77)
    pbc = O.make_pbc(box)
    sw = O.make_switch("RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8")
    nl = O.NeighborList(O.NL_SINGLELIST, n, 0, cutoff=1.0, stride=10)
    nl.update(pbc, pos, fast=True)
    v, d, w, _ = O.coordination(nl, pbc, True, sw, pos)
    want = np.concatenate([d.ravel(), w.ravel(), [v]])
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    sl = np.load(tmp_path / "slices.npy")
    assert sl[0, 0] == 0 and sl[0, 0] + sl[0, 1] == sl[1, 0] and sl[1, 0] + sl[1, 1] == n


@pytest.mark.parametrize("n,world", [(10, 1), (10, 2), (10, 3), (10, 4), (1000003, 8), (7, 8), (1, 2)])
def test_shard_range_partitions(n, world):
    spans = [P.shard_range(n, r, world) for r in range(world)]
    assert sum(c for _, c in spans) == n
    pos = 0
    for lo, c in spans:
        assert lo == min(n, pos) and c >= 0
        pos += c
    chunk = (n + world - 1) // world
    assert all(c == chunk for _, c in spans[:-1] if c) or n < world
