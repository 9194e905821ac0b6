"""x-slab decomposition arithmetic shared by the host mirror, bench.py and the tests.

The same formulas are implemented on the device side in csrc/comm.cuh (comm_split_range and the transpose pack / unpack
index maps); tests/test_slab_gloo.py exercises them across two processes (gloo) against single-process numpy results.
"""
from __future__ import annotations

import numpy as np


def split_range(n: int, n_ranks: int, rank: int):
    """Block partition of range(n): the first n % P ranks get one extra element. Returns (start, count)."""
    base, rem = divmod(n, n_ranks)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def slab(Nx: int, n_ranks: int, rank: int):
    """Rank-local x-range of the slab decomposition (Nx must divide evenly, as the library requires)."""
    if Nx % n_ranks:
        raise ValueError("Nx must be divisible by the number of ranks")
    nx = Nx // n_ranks
    return rank * nx, nx


def pack_forward(W_slab: np.ndarray, n_ranks: int):
    """W_slab[k, ky, i_local] → list of per-peer send blocks [k, ky in peer's range, i_local] (comm.cuh transpose_pack_fwd)."""
    nky = W_slab.shape[1]
    out = []
    for p in range(n_ranks):
        s, c = split_range(nky, n_ranks, p)
        out.append(np.ascontiguousarray(W_slab[:, s:s + c, :]))
    return out


def unpack_forward(blocks, nx: int):
    """Blocks received from every peer ([k, ky_local, i_of_peer]) → W2[k, ky_local, kx_global] (transpose_unpack_fwd)."""
    return np.concatenate(blocks, axis=2)


def pack_backward(W2: np.ndarray, n_ranks: int):
    nx = W2.shape[2] // n_ranks
    return [np.ascontiguousarray(W2[:, :, p * nx:(p + 1) * nx]) for p in range(n_ranks)]


def unpack_backward(blocks):
    return np.concatenate(blocks, axis=1)
