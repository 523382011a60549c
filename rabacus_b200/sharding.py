"""Index conventions of the multi-GPU decompositions (pure Python / numpy).

The CUDA kernels ``k_pack_cells`` / ``k_unpack_cells`` (csrc/rb_kernels.cuh)
implement exactly these two functions on the device; they are kept here as the
executable specification that the CPU (gloo) tests exercise.

Cells: the swept cells ``0 .. ncell-1`` are dealt cyclically, rank ``g`` owning
``il = g, g + G, g + 2G, ...`` (ray cost grows smoothly with depth, so a cyclic
deal balances it).  After each sweep every rank contributes ``[n_fields, n_local]``
and receives ``[G, n_fields, n_local]``; cell ``il = j*G + g`` comes from rank ``g``.

Over peer memory the deal is by BLOCKS of 32 consecutive cells where the grid allows it
(``ncell % (32 G) == 0``): rank ``g`` owns blocks ``g, g + G, ...``, local cell ``j`` is
``il = ((j // 32) G + g) 32 + j % 32``.  The 32 lanes of a ray-walk warp then hold 32
consecutive shells whatever ``G`` (cyclic cells at ``G = 8`` spread a warp over 256 shells
and cost 2.5x the instructions per ray); the ray cost still balances because the blocks
interleave.  Both deals are written as a ``(start, stride)`` pair, ``stride = +G`` cyclic,
``stride = -G`` blocks (``cell_of`` in csrc/rb_common.cuh, ``rb_cell_of`` in the C ABI).

Problems (parameter sweeps): contiguous blocks, no exchange at all.
"""
import numpy as np

BLOCK = 32


def cell_of(j, start, stride):
    """Local cell index ``j`` -> cell (the twin of csrc/rb_common.cuh ``cell_of``)."""
    j = np.asarray(j)
    if stride > 0:
        return start + j * stride
    return ((j // BLOCK) * (-stride) + start) * BLOCK + j % BLOCK


def peer_deal(world, ncell):
    """Stride code the peer path uses for ``ncell`` swept cells on ``world`` ranks."""
    return -world if ncell % (BLOCK * world) == 0 else world


def local_cells(rank, world, ncell, deal=None):
    """Cells of ``rank``: cyclic (default) or by the stride code ``deal`` (+-world)."""
    if deal is None or deal > 0:
        return np.arange(rank, ncell, world)
    return cell_of(np.arange(ncell // world), rank, deal)


def pack_cells(fields, rank, world, ncell, deal=None):
    """fields: [n_fields, Nl] -> send buffer [n_fields, n_local]"""
    return np.ascontiguousarray(fields[:, local_cells(rank, world, ncell, deal)])


def unpack_cells(recv, fields, world, ncell, mirror=False, deal=None):
    """recv: [world, n_fields, n_local]; scatter into fields [n_fields, Nl] in place.
    mirror: two-sided slabs also fill the mirrored half."""
    Nl = fields.shape[1]
    for g in range(world):
        idx = local_cells(g, world, ncell, deal)
        fields[:, idx] = recv[g]
        if mirror:
            fields[:, Nl - 1 - idx] = recv[g]
    return fields


def shard_problems(n_problems, rank, world):
    per = (n_problems + world - 1) // world
    lo = min(n_problems, rank * per)
    return lo, min(n_problems, lo + per)


def deal_problems(n_problems, rank, world, seed=0):
    """Shuffled deal of a problem batch: a fixed pseudo-random permutation (same on every
    rank) dealt cyclically; returns this rank's problem indices, ascending.  A parameter
    sweep is ordered along its axes and the sweep count varies smoothly along them, so
    contiguous blocks leave the rank with the slowest corner of the grid behind (C5 with
    find_Teq on 8 GPUs: 0.84 s for the slowest block against 0.61 s for the fastest), and
    a plain cyclic deal aliases with an axis whose length is a multiple of the world size
    (z has 16 values: ranks 0..7 would only ever see two redshifts each).  No
    communication in any case."""
    perm = np.random.default_rng(seed).permutation(n_problems)
    return np.sort(perm[rank::world])
