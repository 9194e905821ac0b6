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


# ---- peer-blocked flat layouts of the two spectral arrays (csrc/poisson.cuh w_index / w2_index, api.cu transpose_chunk_dma) ----------
# W  (x-slab side,  nx x nky x Nz):      block p = the ky range owned by rank p  -> [p][k][ky - start_p][i]
# W2 (transposed,   Nx x nky_loc x Nz):  block p = the x columns of rank p       -> [p][k][ky_loc][i_p]
# Both are flat arrays of complex numbers; a z chunk [k0, k1) of any block is ONE contiguous run, which is what lets a transpose be
# n_ranks contiguous copies (cudaMemcpyAsync over NVLink, or ncclSend / ncclRecv) per chunk.

def w_offset(nky: int, n_ranks: int, nx: int, Nz: int, k: int, ky: int, i: int) -> int:
    """Offset of (k, ky, i) in a rank's flat W (poisson.cuh w_index, P > 1 branch)."""
    for p in range(n_ranks):
        st, cnt = split_range(nky, n_ranks, p)
        if st <= ky < st + cnt:
            return st * nx * Nz + (k * cnt + (ky - st)) * nx + i
    raise IndexError(ky)


def w2_offset(nky_loc: int, nx: int, Nz: int, k: int, ky_loc: int, kx: int) -> int:
    """Offset of (k, ky_loc, kx_global) in a rank's flat W2 (poisson.cuh w2_index)."""
    p, i = divmod(kx, nx)
    return ((p * Nz + k) * nky_loc + ky_loc) * nx + i


def transpose_chunk_plan(forward: bool, nky: int, n_ranks: int, rank: int, nx: int, Nz: int, k0: int, k1: int):
    """The copies of one z chunk of a transpose as rank `rank` issues them (api.cu transpose_chunk_dma): a list of
    (peer, offset in the PEER's source array, offset in MY destination array, count). forward: my W2 block p <- rank p's W block `rank`;
    backward: my W block p <- rank p's W2 block `rank`."""
    st_me, cnt_me = split_range(nky, n_ranks, rank)
    plan = []
    for q in range(n_ranks):
        p = (rank + q) % n_ranks
        st_p, cnt_p = split_range(nky, n_ranks, p)
        if forward:
            count = (k1 - k0) * cnt_me * nx
            src = st_me * nx * Nz + k0 * cnt_me * nx
            dst = p * Nz * cnt_me * nx + k0 * cnt_me * nx
        else:
            count = (k1 - k0) * cnt_p * nx
            src = rank * Nz * cnt_p * nx + k0 * cnt_p * nx
            dst = st_p * nx * Nz + k0 * cnt_p * nx
        if count:
            plan.append((p, src, dst, count))
    return plan


def packed_face_index(nf: int, w: int, Ny: int, Nz: int, side: int, f: int, k: int, j: int, c: int, side0: int = 0) -> int:
    """Offset of column c (of w) of field f at (j, k) in a packed x-face region (comm.cuh pack_faces_both): [side - side0][f][k][j][c]."""
    per_field = w * Ny * Nz
    return (side - side0) * per_field * nf + f * per_field + (k * Ny + j) * w + c
