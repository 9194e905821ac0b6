// comm.cuh — x-slab decomposition across the GPUs of one box: NCCL (NVLink 5 / NVSwitch) point-to-point groups.
//
// The reference has no multi-GPU path at all; this is the B200-side design:
//   * WENO ghost cells: the BZ_HALO-wide x-faces of the prognostics travel to the two periodic neighbours in ONE
//     ncclGroup (send left + send right + 2 recv) between a pack and an unpack kernel;
//   * distributed FFT: after the local y transform the spectrum W[k][ky][i_local] is re-partitioned to
//     W2[k][ky_local][kx_global] by a grouped all-to-all (ncclSend/ncclRecv per peer), so that the x transform and
//     the z Thomas solve are local; the inverse runs the same exchange backwards.
// libnccl is dlopen'ed on first multi-rank use (torch's bundled libnccl.so.2 if already loaded, else the system
// one), so a single-GPU context has no NCCL dependency. One process per GPU; rank r's periodic neighbours are r±1.
#pragma once
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "poisson.cuh"
#include "aux_kernels.cuh"

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_t;
enum { NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2 };      // ncclDataType_t / ncclRedOp_t values (nccl.h)  BZ_KEEP_F64

struct NcclApi {
    void* lib = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_t, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

struct Comm {
    int n_ranks = 1, rank = 0;
    NcclApi api;
    ncclComm_t comm = nullptr;
    double* halo_send = nullptr;    // [2][nf][HX][Ny][Nz]
    double* halo_recv = nullptr;
    size_t halo_cap = 0;
    // Peer memory (CUDA IPC): every rank maps the other ranks' arenas; halo ghosts and the FFT transposes are then peer
    // LOADS issued by the consuming kernels over NVLink, ordered by a tiny NCCL all-reduce used as a stream barrier.
    int p2p = 0;
    char* peer_base[8] = {};        // arena base of every rank in THIS process' address space (own arena at [rank])
    // Flag barriers in peer memory: every arena ends with a page of 64-bit flags, flag[channel][source rank]. A rank ARRIVES by storing the
    // channel's next epoch into its slot on every peer of the barrier (st.release.sys over NVLink) and WAITS until its own copy of every
    // peer's slot has reached that epoch (ld.acquire.sys). One ~3 us kernel instead of a one-element ncclAllReduce; channels keep the
    // barriers of the two streams apart. use_flags = 0 (BZ_NCCL_BARRIER=1) falls back to the NCCL all-reduce.
    long long off_flags = 0;
    unsigned long long epoch[8] = {};
    int use_flags = 1;
    // Packed x faces in peer memory: the owner gathers its edge columns into a contiguous region of its arena, the neighbours read that
    // region with fully coalesced peer loads and scatter it into their ghost cells. (Reading the 32-byte column segments straight out of
    // the neighbour's padded field was 3 x slower over NVLink at 8 slabs.) Four regions — ρu face, φ, scalars, all / momentum — so that
    // two uses of one region are always separated by at least one full barrier. BZ_DIRECT_PULL=1 selects the old direct reads.
    long long off_pack = 0;
    size_t pack_start[4] = {}, pack_bytes = 0;
    int packed_pull = 1;
    double* token = nullptr;
    char err[256] = {};
};

#define NCCL_TRY(cm, call)                                                                                  \
    do {                                                                                                    \
        int r_ = (call);                                                                                    \
        if (r_ != 0) {                                                                                      \
            snprintf((cm).err, 256, "%s: %s", #call, (cm).api.GetErrorString ? (cm).api.GetErrorString(r_) : "nccl error"); \
            return BZ_ERR_NCCL;                                                                             \
        }                                                                                                   \
    } while (0)

static inline void comm_split_range(int n, int P, int r, int* start, int* count) {
    int base = n / P, rem = n % P;
    *count = base + (r < rem ? 1 : 0);
    *start = r * base + (r < rem ? r : rem);
}
static inline void comm_split_ky(const Comm& cm, int nky, int* ky0, int* nky_loc) { comm_split_range(nky, cm.n_ranks, cm.rank, ky0, nky_loc); }

static void* comm_dlopen_nccl() {
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    void* lib = nullptr;
    for (int n = 0; names[n] && !lib; ++n) lib = dlopen(names[n], RTLD_NOW | RTLD_GLOBAL);
    return lib;
}

// ncclGetUniqueId for rank 0; the caller broadcasts the 128 bytes to the other ranks (any transport).
static int comm_unique_id(uint8_t* out, char* err) {
    void* lib = comm_dlopen_nccl();
    if (!lib) { snprintf(err, 256, "cannot dlopen libnccl.so.2: %s", dlerror()); return BZ_ERR_NCCL; }
    int (*get)(ncclUniqueId_t*) = nullptr;
    *(void**)(&get) = dlsym(lib, "ncclGetUniqueId");
    ncclUniqueId_t id;
    if (!get || get(&id) != 0) { snprintf(err, 256, "ncclGetUniqueId failed"); return BZ_ERR_NCCL; }
    memcpy(out, id.internal, 128);
    return BZ_OK;
}

static int comm_init(Comm& cm, const bz_config* cfg, cudaStream_t) {
    cm.n_ranks = cfg->n_ranks < 1 ? 1 : cfg->n_ranks;
    cm.rank = cfg->rank;
    if (cm.n_ranks == 1) return BZ_OK;
    cm.api.lib = comm_dlopen_nccl();
    if (!cm.api.lib) { snprintf(cm.err, 256, "cannot dlopen libnccl.so.2: %s", dlerror()); return BZ_ERR_NCCL; }
#define SYM(field, name) *(void**)(&cm.api.field) = dlsym(cm.api.lib, name); if (!cm.api.field) { snprintf(cm.err, 256, "missing symbol %s", name); return BZ_ERR_NCCL; }
    SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    ncclUniqueId_t id;
    memcpy(id.internal, cfg->nccl_unique_id, 128);
    NCCL_TRY(cm, cm.api.CommInitRank(&cm.comm, cm.n_ranks, id, cm.rank));
    return BZ_OK;
}

static void comm_destroy(Comm& cm) {
    if (cm.comm && cm.api.CommDestroy) cm.api.CommDestroy(cm.comm);
    cm.comm = nullptr;
    for (int p = 0; p < cm.n_ranks && cm.p2p; ++p) if (p != cm.rank && cm.peer_base[p]) cudaIpcCloseMemHandle(cm.peer_base[p]);
    cm.p2p = 0;
    cudaFree(cm.halo_send); cudaFree(cm.halo_recv); cudaFree(cm.token);
    cm.halo_send = cm.halo_recv = nullptr; cm.token = nullptr;
}

static int comm_alloc_buffers(Comm& cm, const Layout& L, const PoissonGeom&, int64_t* bytes) {
    if (cm.n_ranks == 1) return BZ_OK;
    cm.halo_cap = (size_t)2 * (NPROG + 1) * L.HX * L.Ny * L.Nz;
    if (cudaMalloc(&cm.halo_send, cm.halo_cap * sizeof(double)) || cudaMalloc(&cm.halo_recv, cm.halo_cap * sizeof(double)) || cudaMalloc(&cm.token, 64) || cudaMemset(cm.token, 0, 64)) {
        snprintf(cm.err, 256, "cudaMalloc of communication buffers failed"); return BZ_ERR_NOMEM;
    }
    *bytes += (int64_t)(2 * cm.halo_cap * sizeof(double));
    return BZ_OK;
}

// Periodic x-halo exchange of F.n fields: my right-most HX columns go to the right neighbour's left ghosts and vice versa.
static int comm_exchange_x_halos(Comm& cm, const Layout& L, const FieldSet& F, cudaStream_t s, int64_t* launches) {
    const int P = cm.n_ranks, left = (cm.rank + P - 1) % P, right = (cm.rank + 1) % P;
    const size_t face = (size_t)F.n * L.HX * L.Ny * L.Nz;
    const int blocks = (int)((face + 255) / 256) > 148 * 8 ? 148 * 8 : (int)((face + 255) / 256);
    pack_x_faces<<<blocks, 256, 0, s>>>(L, F, 0, L.HX, cm.halo_send);                       // my left-most interior columns → left neighbour
    pack_x_faces<<<blocks, 256, 0, s>>>(L, F, L.nx - L.HX, L.HX, cm.halo_send + face);      // my right-most interior columns → right neighbour
    NCCL_TRY(cm, cm.api.GroupStart());
    NCCL_TRY(cm, cm.api.Send(cm.halo_send, face, NCCL_FLOAT64, left, cm.comm, s));
    NCCL_TRY(cm, cm.api.Send(cm.halo_send + face, face, NCCL_FLOAT64, right, cm.comm, s));
    NCCL_TRY(cm, cm.api.Recv(cm.halo_recv, face, NCCL_FLOAT64, right, cm.comm, s));         // right neighbour's left-most columns → my right ghosts
    NCCL_TRY(cm, cm.api.Recv(cm.halo_recv + face, face, NCCL_FLOAT64, left, cm.comm, s));   // left neighbour's right-most columns → my left ghosts
    NCCL_TRY(cm, cm.api.GroupEnd());
    unpack_x_faces<<<blocks, 256, 0, s>>>(L, F, L.nx, L.HX, cm.halo_recv);
    unpack_x_faces<<<blocks, 256, 0, s>>>(L, F, -L.HX, L.HX, cm.halo_recv + face);
    *launches += 4;
    return BZ_OK;
}

// ρu at the first ghost face (i = nx) = the right neighbour's first column: one column, one direction.
static int comm_exchange_u_face(Comm& cm, const Layout& L, double* ru, cudaStream_t s, int64_t* launches) {
    const int P = cm.n_ranks, left = (cm.rank + P - 1) % P, right = (cm.rank + 1) % P;
    FieldSet F; F.n = 1; F.f[0] = ru;
    const size_t face = (size_t)L.Ny * L.Nz;
    const int blocks = (int)((face + 255) / 256) > 148 * 8 ? 148 * 8 : (int)((face + 255) / 256);
    pack_x_faces<<<blocks, 256, 0, s>>>(L, F, 0, 1, cm.halo_send);
    NCCL_TRY(cm, cm.api.GroupStart());
    NCCL_TRY(cm, cm.api.Send(cm.halo_send, face, NCCL_FLOAT64, left, cm.comm, s));
    NCCL_TRY(cm, cm.api.Recv(cm.halo_recv, face, NCCL_FLOAT64, right, cm.comm, s));
    NCCL_TRY(cm, cm.api.GroupEnd());
    unpack_x_faces<<<blocks, 256, 0, s>>>(L, F, L.nx, 1, cm.halo_recv);
    *launches += 2;
    return BZ_OK;
}

// All-to-all of the distributed FFT. Both spectral arrays are kept in peer-blocked layouts (poisson.cuh), so every
// ncclSend / ncclRecv moves one contiguous block and no pack / unpack kernel is needed.
static int comm_alltoall(Comm& cm, const double2* send, double2* recv, int nx, const PoissonGeom& G, bool forward, cudaStream_t s) {
    const int P = cm.n_ranks;
    NCCL_TRY(cm, cm.api.GroupStart());
    for (int p = 0; p < P; ++p) {
        int st, cnt;
        comm_split_range(G.nky, P, p, &st, &cnt);
        // x-slab side: block p = peer p's ky range of my columns; transposed side: block p = my ky range of peer p's columns
        size_t slab_off = (size_t)st * nx * G.Nz, slab_n = (size_t)cnt * nx * G.Nz;
        size_t tr_off = (size_t)p * nx * G.nky_loc * G.Nz, tr_n = (size_t)nx * G.nky_loc * G.Nz;
        if (forward) {
            if (slab_n) NCCL_TRY(cm, cm.api.Send(send + slab_off, slab_n * 2, NCCL_FLOAT64, p, cm.comm, s));
            if (tr_n) NCCL_TRY(cm, cm.api.Recv(recv + tr_off, tr_n * 2, NCCL_FLOAT64, p, cm.comm, s));
        } else {
            if (tr_n) NCCL_TRY(cm, cm.api.Send(send + tr_off, tr_n * 2, NCCL_FLOAT64, p, cm.comm, s));
            if (slab_n) NCCL_TRY(cm, cm.api.Recv(recv + slab_off, slab_n * 2, NCCL_FLOAT64, p, cm.comm, s));
        }
    }
    NCCL_TRY(cm, cm.api.GroupEnd());
    return BZ_OK;
}

static int comm_transpose_forward(Comm& cm, const double2* W, double2* W2, int nx, const PoissonGeom& G, cudaStream_t s, int64_t*) {
    return comm_alltoall(cm, W, W2, nx, G, true, s);
}
static int comm_transpose_backward(Comm& cm, const double2* W2, double2* W, int nx, const PoissonGeom& G, cudaStream_t s, int64_t*) {
    return comm_alltoall(cm, W2, W, nx, G, false, s);
}

#define BZ_FLAG_CHANNELS 8
#define BZ_FLAG_TIMEOUT_SLOT (BZ_FLAG_CHANNELS * 8)      // set to 1 by a barrier that gave up waiting (a peer died): checked by bz_synchronize

struct PeerFlagPtrs { unsigned long long* base[8]; };

__global__ void peer_barrier_kernel(PeerFlagPtrs flags, int rank, int n_ranks, unsigned mask, int channel, unsigned long long epoch) {
    const int p = threadIdx.x;
    if (p >= n_ranks || p == rank || !((mask >> p) & 1u)) return;
    __threadfence_system();                                           // everything this stream wrote before the barrier is visible to the peers first
    unsigned long long* remote = flags.base[p] + channel * 8 + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(epoch) : "memory");
    const unsigned long long* mine = flags.base[rank] + channel * 8 + p;
    volatile unsigned long long* dead = flags.base[rank] + BZ_FLAG_TIMEOUT_SLOT;
    if (*dead) return;                                                // an earlier barrier gave up: drain the queue, bz_synchronize reports it
    const long long t0 = clock64();
    unsigned long long seen;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
        if (seen >= epoch) return;
        __nanosleep(100);
    } while (clock64() - t0 < 10000000000LL);                         // ~5 s at 2 GHz: never hang the GPU on a dead peer
    *dead = 1ull;
}

// Stream barrier across ranks (all of them, or the ranks in `mask`): when it completes on this rank's stream, every participating rank's
// earlier work on the stream it issued ITS barrier on is complete. Masks must be symmetric and all ranks must issue the same barrier
// sequence per channel.
static int comm_barrier(Comm& cm, cudaStream_t s, int channel = 0, unsigned mask = 0xffu) {
    if (cm.p2p && cm.use_flags) {
        PeerFlagPtrs F;
        for (int p = 0; p < 8; ++p) F.base[p] = cm.peer_base[p] ? reinterpret_cast<unsigned long long*>(cm.peer_base[p] + cm.off_flags) : nullptr;
        peer_barrier_kernel<<<1, 32, 0, s>>>(F, cm.rank, cm.n_ranks, mask, channel, ++cm.epoch[channel]);
        if (cudaGetLastError() != cudaSuccess) { snprintf(cm.err, 256, "peer barrier launch failed"); return BZ_ERR_CUDA; }
        return BZ_OK;
    }
    NCCL_TRY(cm, cm.api.AllReduce(cm.token, cm.token, 1, NCCL_FLOAT64, NCCL_MAX, cm.comm, s));
    return BZ_OK;
}
static unsigned comm_neighbour_mask(const Comm& cm) {
    const int P = cm.n_ranks;
    return (1u << ((cm.rank + P - 1) % P)) | (1u << ((cm.rank + 1) % P));
}

static int comm_ipc_export(void* arena, uint8_t* out64, char* err) {
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, arena);
    if (e != cudaSuccess) { snprintf(err, 256, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return BZ_ERR_CUDA; }
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out64, &h, 64);
    return BZ_OK;
}

// handles: n_ranks × 64 bytes, rank-ordered (all_gather of bz_ipc_export).
static int comm_ipc_attach(Comm& cm, void* own_arena, const uint8_t* handles) {
    if (cm.n_ranks == 1) return BZ_OK;
    if (cm.n_ranks > 8) { snprintf(cm.err, 256, "peer memory path supports up to 8 ranks"); return BZ_ERR_UNSUPPORTED; }
    for (int p = 0; p < cm.n_ranks; ++p) {
        if (p == cm.rank) { cm.peer_base[p] = (char*)own_arena; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)64 * p, 64);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { snprintf(cm.err, 256, "cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e)); return BZ_ERR_CUDA; }
        cm.peer_base[p] = (char*)ptr;
    }
    cm.p2p = 1;
    if (const char* e = getenv("BZ_NCCL_BARRIER")) cm.use_flags = atoi(e) ? 0 : 1;
    if (const char* e = getenv("BZ_DIRECT_PULL")) cm.packed_pull = atoi(e) ? 0 : 1;
    return BZ_OK;
}

// Ghost columns pulled straight from the neighbours' interiors (peer loads). mode 0: both sides, HX columns;
// mode 1: only the first ghost column on the right (ρu at i = nx for the divergence).
__global__ void halo_pull_x(Layout L, FieldSet F, const char* my_base, const char* left_base, const char* right_base, int mode) {
    const int ncol = mode == 0 ? 2 * L.HX : 1;
    const long long per_field = (long long)ncol * L.Ny * L.Nz;
    const long long total = per_field * F.n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int f = (int)(e / per_field); long long r = e % per_field;
        int c = (int)(r % ncol), j = (int)((r / ncol) % L.Ny), k = (int)(r / ((long long)ncol * L.Ny));
        int i_dst, i_src; const char* peer;
        if (mode == 0 && c < L.HX) { i_dst = c - L.HX; i_src = L.nx - L.HX + c; peer = left_base; }
        else { int cc = mode == 0 ? c - L.HX : 0; i_dst = L.nx + cc; i_src = cc; peer = right_base; }
        const double* src = reinterpret_cast<const double*>(peer + (reinterpret_cast<const char*>(F.f[f]) - my_base));
        F.f[f][lidx(L, i_dst, j, k)] = __ldcv(src + lidx(L, i_src, j, k));
    }
}

// Packed variant, owner side: side 0 = my first w columns (the LEFT neighbour's right ghosts), side 1 = my last w columns (the RIGHT
// neighbour's left ghosts); layout [side - side0][field][k][j][c], c fastest; sides side0 .. side0 + nsides - 1 are packed.
__global__ void pack_faces_both(Layout L, FieldSet F, int w, int nsides, int side0, double* __restrict__ buf) {
    const long long per_field = (long long)w * L.Ny * L.Nz, per_side = per_field * F.n, total = per_side * nsides;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(e / per_side); const long long q = e - s * per_side;
        const int f = (int)(q / per_field); const long long r = q - f * per_field;
        const int c = (int)(r % w), j = (int)((r / w) % L.Ny), k = (int)(r / ((long long)w * L.Ny));
        buf[e] = F.f[f][lidx(L, (side0 + s == 0 ? 0 : L.nx - w) + c, j, k)];
    }
}
// consumer side: right ghosts from the right neighbour's side 0, left ghosts from the left neighbour's side 1
__global__ void unpack_faces_both(Layout L, FieldSet F, int w, int nsides, int side0, const double* __restrict__ left_buf, const double* __restrict__ right_buf) {
    const long long per_field = (long long)w * L.Ny * L.Nz, per_side = per_field * F.n, total = per_side * nsides;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(e / per_side); const long long q = e - s * per_side;
        const int f = (int)(q / per_field); const long long r = q - f * per_field;
        const int c = (int)(r % w), j = (int)((r / w) % L.Ny), k = (int)(r / ((long long)w * L.Ny));
        const int side = side0 + s;
        const double v = side == 0 ? __ldcv(right_buf + e) : __ldcv(left_buf + e);
        F.f[f][lidx(L, (side == 0 ? L.nx : -w) + c, j, k)] = v;
    }
}

// mode 0: HX ghost columns on both sides; mode 1: the first ghost column on the right only (ρu at i = nx for the divergence);
// mode 2: the first ghost column on the left only (φ at i = -1 for the projection).
// region: 0 ρu face, 1 φ, 2 scalars, 3 all five / momentum (Comm::pack_start)
static int comm_pull_x_halos(Comm& cm, const Layout& L, const FieldSet& F, int mode, cudaStream_t s, int64_t* launches, int channel = 0, int region = 3) {
    const int P = cm.n_ranks, left = (cm.rank + P - 1) % P, right = (cm.rank + 1) % P;
    if (!cm.packed_pull && mode == 2) mode = 0;                        // the direct reads know modes 0 and 1
    const int w = mode == 0 ? L.HX : 1, nsides = mode == 0 ? 2 : 1, side0 = mode == 2 ? 1 : 0;
    const long long total = (long long)F.n * nsides * w * L.Ny * L.Nz;
    const int blocks = (int)((total + 255) / 256) > 148 * 8 ? 148 * 8 : (int)((total + 255) / 256);
    if (cm.packed_pull) {
        double* mine = reinterpret_cast<double*>(cm.peer_base[cm.rank] + cm.off_pack) + cm.pack_start[region];
        pack_faces_both<<<blocks, 256, 0, s>>>(L, F, w, nsides, side0, mine);
        *launches += 1;
    }
    int rc = comm_barrier(cm, s, channel, comm_neighbour_mask(cm));   // only the two neighbours' fields are read
    if (rc) return rc;
    if (cm.packed_pull) {
        const double* lbuf = reinterpret_cast<const double*>(cm.peer_base[left] + cm.off_pack) + cm.pack_start[region];
        const double* rbuf = reinterpret_cast<const double*>(cm.peer_base[right] + cm.off_pack) + cm.pack_start[region];
        unpack_faces_both<<<blocks, 256, 0, s>>>(L, F, w, nsides, side0, lbuf, rbuf);
    } else {
        halo_pull_x<<<blocks, 256, 0, s>>>(L, F, cm.peer_base[cm.rank], cm.peer_base[left], cm.peer_base[right], mode);
    }
    *launches += 1;
    return BZ_OK;
}

static int comm_allreduce_sum_device(Comm& cm, double* dev, size_t n, cudaStream_t s) {
    NCCL_TRY(cm, cm.api.AllReduce(dev, dev, n, NCCL_FLOAT64, NCCL_SUM, cm.comm, s));
    return BZ_OK;
}

static int comm_allreduce_max(Comm& cm, double* host_value, double* dev_scalar, cudaStream_t s) {
    if (cudaMemcpyAsync(dev_scalar, host_value, sizeof(double), cudaMemcpyHostToDevice, s)) return BZ_ERR_CUDA;
    NCCL_TRY(cm, cm.api.AllReduce(dev_scalar, dev_scalar, 1, NCCL_FLOAT64, NCCL_MAX, cm.comm, s));
    if (cudaMemcpyAsync(host_value, dev_scalar, sizeof(double), cudaMemcpyDeviceToHost, s)) return BZ_ERR_CUDA;
    if (cudaStreamSynchronize(s)) return BZ_ERR_CUDA;
    return BZ_OK;
}
