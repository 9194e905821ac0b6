"""world_size-2 gloo tests (CPU) of the host-side slab logic: decomposition arithmetic, slab-local initial conditions, the
periodic halo exchange pattern and the distributed-FFT transposes, against single-process numpy results."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from breeze_b200 import slab
        Nx, Ny, Nz, H = 16, 8, 6, 4
        rng = np.random.default_rng(0)
        full = rng.standard_normal((Nz, Ny, Nx))
        i0, nx = slab.slab(Nx, world, rank)
        mine = full[:, :, i0:i0 + nx].copy()

        # (1) slabs tile the domain
        gathered = [torch.zeros(Nz, Ny, nx, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine))
        assert np.array_equal(np.concatenate([g.numpy() for g in gathered], axis=2), full)

        # (2) periodic halo exchange: my left-most columns → left neighbour's right ghosts, right-most → right neighbour's left ghosts
        left, right = (rank - 1) % world, (rank + 1) % world
        send_l, send_r = torch.from_numpy(mine[:, :, :H].copy()), torch.from_numpy(mine[:, :, -H:].copy())
        recv_from_right, recv_from_left = torch.zeros_like(send_l), torch.zeros_like(send_r)
        ops = [dist.P2POp(dist.isend, send_l, left), dist.P2POp(dist.isend, send_r, right),
               dist.P2POp(dist.irecv, recv_from_right, right), dist.P2POp(dist.irecv, recv_from_left, left)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        padded = np.concatenate([recv_from_left.numpy(), mine, recv_from_right.numpy()], axis=2)
        idx = (np.arange(i0 - H, i0 + nx + H)) % Nx
        assert np.array_equal(padded, full[:, :, idx])

        # (3) distributed FFT: y transform local → all-to-all (block partition of ky) → x transform == global fft2 (half spectrum)
        Wy = np.fft.rfft(mine, axis=1)                                   # [k, ky, i_local]
        nky = Wy.shape[1]
        send = slab.pack_forward(Wy, world)
        ky0, nky_loc = slab.split_range(nky, world, rank)
        recv = [torch.zeros(Nz, nky_loc, nx, dtype=torch.complex128) for _ in range(world)]
        dist.all_to_all(recv, [torch.from_numpy(b) for b in send]) if dist.get_backend() != "gloo" else None
        if dist.get_backend() == "gloo":                                 # gloo has no all_to_all: grouped send/recv like the NCCL path
            ops = [dist.P2POp(dist.isend, torch.from_numpy(send[p]), p) for p in range(world) if p != rank]
            ops += [dist.P2POp(dist.irecv, recv[p], p) for p in range(world) if p != rank]
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            recv[rank] = torch.from_numpy(send[rank])
        W2 = slab.unpack_forward([r.numpy() for r in recv], nx)          # [k, ky_local, kx_global]
        W2 = np.fft.fft(W2, axis=2)
        ref = np.fft.fft(np.fft.rfft(full, axis=1), axis=2)[:, ky0:ky0 + nky_loc, :]
        assert np.allclose(W2, ref, rtol=1e-12, atol=1e-12)

        # (4) and back
        back = slab.pack_backward(np.fft.ifft(W2, axis=2), world)
        s_counts = [slab.split_range(nky, world, p)[1] for p in range(world)]
        recvb = [torch.zeros(Nz, s_counts[p], nx, dtype=torch.complex128) for p in range(world)]
        ops = [dist.P2POp(dist.isend, torch.from_numpy(back[p]), p) for p in range(world) if p != rank]
        ops += [dist.P2POp(dist.irecv, recvb[p], p) for p in range(world) if p != rank]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        recvb[rank] = torch.from_numpy(back[rank])
        Wy_back = slab.unpack_backward([r.numpy() for r in recvb])
        assert np.allclose(np.fft.irfft(Wy_back, n=Ny, axis=1), mine, rtol=1e-12, atol=1e-12)
        # (5) horizontal means of the subsidence forcing (column_sums + all-reduce across slabs, compute_forcing! of
        #     src/Forcings/subsidence_forcing.jl:137-141): per-level sums over my slab, summed over ranks == global sums
        sums = torch.from_numpy(mine.sum(axis=(1, 2)).copy())
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        assert np.allclose(sums.numpy() / (Nx * Ny), full.mean(axis=(1, 2)), rtol=1e-13, atol=1e-15)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_slab_logic_world_size_2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in results), results


def _flat_worker(rank, world, port, q):
    """The peer-blocked flat layouts and the chunked transposes exactly as api.cu issues them (transpose_chunk_dma): every rank exposes its
    flat W / W2 to the others (all_gather stands in for the mapped peer arenas) and PULLS the contiguous runs of the plan."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from breeze_b200 import slab
        Nx, Ny, Nz, chunks = 24, 12, 7, 3                                # nky = 7: uneven ky blocks (4 + 3 over two ranks, 3 + 2 + 2 over three)
        rng = np.random.default_rng(1)
        full = rng.standard_normal((Nz, Ny, Nx))
        i0, nx = slab.slab(Nx, world, rank)
        Wy = np.fft.rfft(full[:, :, i0:i0 + nx], axis=1)                 # [k, ky, i_local]
        nky = Wy.shape[1]
        ky0, nky_loc = slab.split_range(nky, world, rank)
        # my flat W in the peer-blocked layout
        W = np.zeros(nx * nky * Nz, dtype=np.complex128)
        for k in range(Nz):
            for ky in range(nky):
                o = slab.w_offset(nky, world, nx, Nz, k, ky, 0)
                W[o:o + nx] = Wy[k, ky]
        def gather(flat, n):
            t = torch.zeros(n, dtype=torch.complex128); t[:flat.size] = torch.from_numpy(flat)
            parts = [torch.zeros(n, dtype=torch.complex128) for _ in range(world)]
            dist.all_gather(parts, t)
            return [p_.numpy() for p_ in parts]
        n_w2_max = Nx * max(slab.split_range(nky, world, r)[1] for r in range(world)) * Nz
        peers_W = gather(W, W.size)
        W2 = np.zeros(Nx * nky_loc * Nz, dtype=np.complex128)
        kper = -(-Nz // chunks)
        for ch in range(chunks):                                         # forward transpose, chunk by chunk
            k0, k1 = ch * kper, min(Nz, (ch + 1) * kper)
            for peer, src, dst, count in slab.transpose_chunk_plan(True, nky, world, rank, nx, Nz, k0, k1):
                W2[dst:dst + count] = peers_W[peer][src:src + count]
        ref = np.fft.rfft(full, axis=1)                                  # [k, ky, x_global]
        for k in range(Nz):
            for kyl in range(nky_loc):
                for kx in range(Nx):
                    assert W2[slab.w2_offset(nky_loc, nx, Nz, k, kyl, kx)] == ref[k, ky0 + kyl, kx]
        # x transform in the transposed layout, then the backward transpose and the inverse y transform reproduce the slab
        W2x = W2.copy()
        for k in range(Nz):
            for kyl in range(nky_loc):
                idx = [slab.w2_offset(nky_loc, nx, Nz, k, kyl, kx) for kx in range(Nx)]
                W2x[idx] = np.fft.ifft(np.fft.fft(W2[idx]))
        peers_W2 = gather(W2x, n_w2_max)
        Wb = np.zeros_like(W)
        for ch in range(chunks):
            k0, k1 = ch * kper, min(Nz, (ch + 1) * kper)
            for peer, src, dst, count in slab.transpose_chunk_plan(False, nky, world, rank, nx, Nz, k0, k1):
                Wb[dst:dst + count] = peers_W2[peer][src:src + count]
        assert np.allclose(Wb, W, rtol=1e-13, atol=1e-13)
        # packed x faces (comm.cuh pack_faces_both / unpack_faces_both): my right ghosts are the right neighbour's side 0
        H, nf = 2, 3
        fields = rng.standard_normal((nf, Nz, Ny, Nx))
        mine = fields[:, :, :, i0:i0 + nx]
        buf = np.zeros(2 * nf * H * Ny * Nz)
        for side in (0, 1):
            for f in range(nf):
                for k in range(Nz):
                    for j in range(Ny):
                        for c in range(H):
                            buf[slab.packed_face_index(nf, H, Ny, Nz, side, f, k, j, c)] = mine[f, k, j, (0 if side == 0 else nx - H) + c]
        t = torch.from_numpy(buf); parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        right, left = parts[(rank + 1) % world].numpy(), parts[(rank - 1) % world].numpy()
        for f in range(nf):
            for k in (0, Nz - 1):
                for j in (0, Ny - 1):
                    for c in range(H):
                        assert right[slab.packed_face_index(nf, H, Ny, Nz, 0, f, k, j, c)] == fields[f, k, j, (i0 + nx + c) % Nx]
                        assert left[slab.packed_face_index(nf, H, Ny, Nz, 1, f, k, j, c)] == fields[f, k, j, (i0 - H + c) % Nx]
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()[-600:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_blocked_layouts_and_chunked_transposes(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_flat_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in results), results


def test_split_range_partitions():
    sys.path.insert(0, ROOT)
    from breeze_b200 import slab
    for n in (1, 5, 9, 257):
        for P in (1, 2, 4, 8):
            parts = [slab.split_range(n, P, r) for r in range(P)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (s0, c0), (s1, _) in zip(parts, parts[1:]):
                assert s0 + c0 == s1
