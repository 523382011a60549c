"""N > 1 paths on CPU: two gloo ranks exercise the exchange protocol of the
cell-sharded Jacobi sweep (pack -> all_gather -> unpack, the same index
convention the CUDA kernels k_pack_cells / k_unpack_cells implement) and the
problem partition of the parameter sweep."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rabacus_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_sweep(fields, cells, sweep):
    """Stand-in for the per-cell update: a deterministic function of the WHOLE
    pre-sweep state (as the ray trace is) applied to the local cells only."""
    glob = fields.sum(axis=1, keepdims=True)
    out = fields.copy()
    out[:, cells] = 0.5 * fields[:, cells] + 1e-3 * glob + 0.01 * (sweep + 1) * (cells + 1)
    return out


def _worker(rank, world, port, Nl, mirror, q, deal=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ncell = Nl // 2 if mirror else Nl
    rng = np.random.default_rng(7)                      # same initial state on every rank
    fields = rng.random((12, Nl))
    if mirror:
        fields[:, Nl - ncell:] = fields[:, :ncell][:, ::-1]
    nloc = ncell // world
    for sweep in range(3):
        new = _fake_sweep(fields, sharding.local_cells(rank, world, ncell, deal), sweep)
        send = torch.from_numpy(sharding.pack_cells(new, rank, world, ncell, deal))
        recv = torch.empty((world, 12, nloc), dtype=torch.float64)
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1))
        sharding.unpack_cells(recv.numpy(), fields, world, ncell, mirror=mirror, deal=deal)
    # every rank must now hold the same full state
    t = torch.from_numpy(fields.copy())
    ref = t.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(t, ref))
    if rank == 0:
        q.put((same, fields))
    else:
        q.put((same, None))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mirror,Nl,deal", [(False, 24, None), (True, 24, None),
                                            (False, 128, -2), (True, 256, -2)])
def test_cell_sharded_exchange_equals_single_rank(mirror, Nl, deal):
    """cyclic deal (NCCL path) and the 32-cell block deal of the peer path"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, Nl, mirror, q, deal))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[0] for r in res)
    got = [r[1] for r in res if r[1] is not None][0]
    # single-rank reference: the same Jacobi update over all cells
    ncell = Nl // 2 if mirror else Nl
    rng = np.random.default_rng(7)
    f = rng.random((12, Nl))
    if mirror:
        f[:, Nl - ncell:] = f[:, :ncell][:, ::-1]
    for sweep in range(3):
        new = _fake_sweep(f, np.arange(ncell), sweep)
        f[:, :ncell] = new[:, :ncell]
        if mirror:
            f[:, Nl - ncell:] = new[:, :ncell][:, ::-1]
    assert np.array_equal(got, f)


def test_block_deal_matches_the_library():
    """sharding.cell_of is the twin of csrc/rb_common.cuh cell_of (exported as rb_cell_of);
    both deals are permutations of the swept cells, and a block-dealt warp of 32 local cells
    holds 32 consecutive shells."""
    import ctypes as C
    from rabacus_b200 import _lib
    L = _lib.lib()
    for world, ncell in ((2, 128), (4, 4096), (8, 4096), (8, 256)):
        assert sharding.peer_deal(world, ncell) == -world
        seen = []
        for g in range(world):
            for stride in (world, -world):
                idx = sharding.local_cells(g, world, ncell, stride)
                got = [L.rb_cell_of(C.c_int(j), C.c_int(g), C.c_int(stride)) for j in range(idx.size)]
                assert list(idx) == got
            blk = sharding.local_cells(g, world, ncell, -world)
            assert np.all(np.diff(blk.reshape(-1, 32), axis=1) == 1)
            seen.append(blk)
        assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(ncell))
    assert sharding.peer_deal(8, 4000) == 8 and sharding.peer_deal(2, 33) == 2   # falls back


def test_index_conventions():
    assert list(sharding.local_cells(1, 4, 10)) == [1, 5, 9]
    f = np.arange(36, dtype=float).reshape(3, 12)
    world, ncell = 3, 12
    recv = np.stack([sharding.pack_cells(f * 2, g, world, ncell) for g in range(world)])
    out = sharding.unpack_cells(recv, np.zeros_like(f), world, ncell)
    assert np.array_equal(out, f * 2)
    parts = [sharding.shard_problems(8192, r, 8) for r in range(8)]
    assert parts[0] == (0, 1024) and parts[-1] == (7168, 8192)
    assert sharding.shard_problems(10, 3, 4) == (9, 10) and sharding.shard_problems(3, 3, 4) == (3, 3)
    deals = [sharding.deal_problems(8192, r, 8) for r in range(8)]
    assert all(d.size == 1024 for d in deals) and all(np.all(np.diff(d) > 0) for d in deals)
    # no aliasing with a 16-long fastest axis: every rank sees every value of i % 16
    assert all(np.unique(d % 16).size == 16 for d in deals)
    assert np.array_equal(np.sort(np.concatenate(deals)), np.arange(8192))
    assert [sharding.deal_problems(10, r, 4).size for r in range(4)] == [3, 3, 2, 2]


def test_deals_are_partitions_for_any_size():
    """property test (hypothesis): for every world size and grid the cyclic deal, the block deal
    where the peer path selects it, and the shuffled problem deal are partitions -- every cell /
    problem is owned by exactly one rank -- and pack / unpack is the identity."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 8), st.integers(1, 700), st.booleans())
    def check(world, ncell, mirror):
        deal = sharding.peer_deal(world, ncell)
        assert deal == (-world if ncell % (32 * world) == 0 else world)
        for d in {world, deal}:
            if d > 0 and ncell < world:
                continue
            own = [sharding.local_cells(g, world, ncell, d) for g in range(world)]
            assert np.array_equal(np.sort(np.concatenate(own)), np.arange(ncell))
            if d < 0 or ncell % world == 0:      # the exchange needs equal shares
                Nl = 2 * ncell if mirror else ncell
                f = np.arange(3 * Nl, dtype=float).reshape(3, Nl)
                if mirror:
                    f[:, ncell:] = f[:, :ncell][:, ::-1]
                recv = np.stack([sharding.pack_cells(f, g, world, ncell, d) for g in range(world)])
                out = sharding.unpack_cells(recv, np.zeros_like(f), world, ncell, mirror=mirror,
                                            deal=d)
                assert np.array_equal(out, f)
        n = ncell * 3 + 1
        parts = [sharding.deal_problems(n, r, world) for r in range(world)]
        assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(n))
        assert max(p.size for p in parts) - min(p.size for p in parts) <= 1
        blocks = [sharding.shard_problems(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))

    check()
