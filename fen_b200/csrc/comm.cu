// comm.cu -- multi-GPU exchange over NVLink peer memory (replaces 2decomp&FFT's update_halo and
// transpose_* and the scalar mpi_allreduce calls of the reference: src/halo.f90:33,
// src/poisson.f90:972-1025, src/navier_stokes.f90:614, src/scalar.f90:194,214).
//
// Transport: every rank owns one "arena" in its HBM -- flags, reduction slots, halo mailboxes and
// the two spectral work arrays of the Poisson solver -- and maps the arenas of all other ranks
// (cudaIpcOpenMemHandle between processes, plain pointers + cudaDeviceEnablePeerAccess inside one
// process).  All data movement is done by this library's own kernels with stores to the mapped
// peer addresses; the big ones (the y<->z transposes) are the epilogues of the FFT / tridiagonal
// kernels in poisson.cu, so the spectral field crosses NVLink straight out of shared memory and
// never takes an extra pack/unpack pass through HBM.
//
// Synchronisation: monotonically increasing epoch flags.  A sender finishes its stores, issues
// __threadfence_system() and release-stores the epoch into the receiver's flag; the receiver spins
// on its own flag with acquire loads in a ONE-BLOCK kernel (never inside a fat grid, so a waiting
// rank cannot starve kernels of another stream on the same device).  Mailboxes and reduction slots
// are double-buffered by epoch parity, which makes the write-after-read hazard impossible without
// a second handshake (a rank can only be one exchange ahead of its neighbour).  Every spin has a
// wall-clock bound (globaltimer); on expiry a host-mapped error word is set and the next
// fen_gpu_synchronize / get_status reports FEN_ERR_COMM instead of hanging the GPU.
#include <unistd.h>

#include <cstdlib>
#include <cstring>

#include "fen_internal.cuh"

namespace fen {

// bound of every flag wait; FEN_GPU_SPIN_LIMIT_MS overrides it (tests use a short one)
static unsigned long long spin_limit_ns() {
    static unsigned long long v = 0;
    if (!v) {
        const char* e = getenv("FEN_GPU_SPIN_LIMIT_MS");
        const long long ms = e ? atoll(e) : 0;
        v = (ms > 0 ? (unsigned long long)ms : 20000ull) * 1000000ull;
    }
    return v;
}
constexpr int kSigPerChannel = FEN_MAX_RANKS;
// CH_A2A_CHUNK + q: chunk q of a chunked z -> y transpose (poisson.cu: solve_blocked signals each chunk of granules
// as soon as its stores are out, so that the y inverse of that chunk overlaps the next chunk's transfer)
enum Channel { CH_HALO = 0, CH_A2A_FWD = 1, CH_A2A_BWD = 2, CH_RED = 3, CH_A2A_CHUNK = 4, CH_COUNT = 4 + FEN_MAX_CHUNKS };
constexpr int kRedWidth = 8;     // doubles per allreduce

struct HandleBlob {              // what fen_gpu_comm_export writes (fen_gpu_comm_handle_bytes() bytes)
    unsigned magic;
    int rank, nranks, device;
    long long pid;
    unsigned long long arena_bytes;
    void* raw;                   // valid inside the exporting process only
    cudaIpcMemHandle_t ipc;
    char host[64];
};
constexpr unsigned kMagic = 0x46454e31u;   // "FEN1"

struct Comm {
    int P = 1, rank = 0;
    char* arena = nullptr;
    size_t arena_bytes = 0;
    char* peer[FEN_MAX_RANKS] = {};
    bool ipc_open[FEN_MAX_RANKS] = {};
    bool connected = false;
    size_t off_sig = 0, off_err = 0, off_red = 0, off_mail = 0, off_C = 0, off_Cz = 0;
    size_t plane = 0;            // doubles per z plane of a field (Layout::sz)
    size_t nC = 0, nCz = 0;      // complex elements
    unsigned long long epoch[CH_COUNT] = {};
    int* h_err = nullptr;        // pinned mirror of the error word
};

// The flag waits assume that a kernel of another stream (another rank's, when several ranks share one device
// in one process) can start while a wait kernel spins.  CUDA's lazy module loading breaks that: the first
// launch of a not-yet-loaded kernel synchronises the context and would queue behind the spinning kernel.  So
// the library asks for eager loading when it is loaded (before the runtime initialises); a caller that has
// already set CUDA_MODULE_LOADING keeps its choice.
__attribute__((constructor)) static void fen_request_eager_loading() { setenv("CUDA_MODULE_LOADING", "EAGER", 0); }

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int spectral_pitch(int nx) { return ((nx / 2 + 1) + 7) / 8 * 8; }
// complex row pitch of the Poisson work arrays of this grid: the half spectrum when x is periodic (r2c), the full
// width when x is a Neumann direction (the DCT variants carry real data in a full-width complex array)
int spectral_pitch_grid(const fen_grid_desc& g) {
    const bool perx = g.bc[0] == FEN_BC_PERIODIC && g.bc[1] == FEN_BC_PERIODIC;
    return perx ? spectral_pitch(g.nx) : (g.nx + 7) / 8 * 8;
}

static void comm_layout(fen_ctx* c, Comm* m) {
    const Layout& L = c->L;
    const int P = c->g.nranks;
    m->plane = (size_t)L.sz;
    const int PC = spectral_pitch_grid(c->g);
    m->nC = (size_t)PC * c->g.ny * L.nzl;
    m->nCz = (size_t)PC * (c->g.ny / P) * c->g.nz;
    size_t o = 0;
    m->off_sig = o; o = align_up(o + sizeof(unsigned long long) * CH_COUNT * kSigPerChannel, 256);
    m->off_err = o; o = align_up(o + 256, 256);
    m->off_red = o; o = align_up(o + sizeof(double) * 2 * FEN_MAX_RANKS * kRedWidth, 256);
    m->off_mail = o; o = align_up(o + sizeof(double) * 2 * 2 * 3 * m->plane, 256);
    m->off_C = o; o = align_up(o + sizeof(double2) * m->nC, 256);
    m->off_Cz = o; o = align_up(o + sizeof(double2) * m->nCz, 256);
    m->arena_bytes = o;
}

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void spin_until(const unsigned long long* flag, unsigned long long want, int* err,
                                           unsigned long long limit_ns) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < want) {
        __nanosleep(64);
        if (global_ns() - t0 > limit_ns) {
            *(volatile int*)err = 1;
            break;
        }
    }
}

struct SigArgs {
    unsigned long long* send_flag[FEN_MAX_RANKS];   // flags in the peers' arenas to raise
    const unsigned long long* wait_flag[FEN_MAX_RANKS];   // flags in my arena to wait for
    int nsend, nwait;
    unsigned long long epoch, limit_ns;
    int* err;
};

// raise `epoch` on the peers, then wait until the peers have raised it here; one block
__global__ void k_sigwait(SigArgs a) {
    const int t = threadIdx.x;
    if (t < a.nsend) {
        __threadfence_system();
        st_release_sys(a.send_flag[t], a.epoch);
    }
    if (t < a.nwait) spin_until(a.wait_flag[t], a.epoch, a.err, a.limit_ns);
    __syncthreads();
}

struct HaloArgs {
    const double* src[2][3];     // [side][comp]: my boundary plane that goes to the neighbour on that side
    double* dst[2][3];           // [side][comp]: the neighbour's mailbox slot (mapped)
    int n;                       // components
    int has[2];                  // neighbour present on the lo / hi side
    size_t plane2;               // plane length in double2 units
};

__global__ void __launch_bounds__(256) k_halo_push(HaloArgs a) {
    const int side = blockIdx.y, m = blockIdx.z;
    if (!a.has[side] || m >= a.n) return;
    const double2* __restrict__ s = reinterpret_cast<const double2*>(a.src[side][m]);
    double2* __restrict__ d = reinterpret_cast<double2*>(a.dst[side][m]);
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < a.plane2; e += (size_t)gridDim.x * blockDim.x)
        d[e] = s[e];
}

struct UnpackArgs {
    const double* src[2][3];     // my mailbox slots [side][comp]
    double* dst[2][3];           // ghost planes k = 0 (side 0) and k = nzl + 1 (side 1)
    int n;
    int has[2];
    size_t plane2;
};

__global__ void __launch_bounds__(256) k_halo_unpack(UnpackArgs a) {
    const int side = blockIdx.y, m = blockIdx.z;
    if (!a.has[side] || m >= a.n) return;
    const double2* __restrict__ s = reinterpret_cast<const double2*>(a.src[side][m]);
    double2* __restrict__ d = reinterpret_cast<double2*>(a.dst[side][m]);
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < a.plane2; e += (size_t)gridDim.x * blockDim.x)
        d[e] = s[e];
}

struct RedArgs {
    double* vals;                                 // in/out, n values
    double* slot_peer[FEN_MAX_RANKS];             // peers' slot[parity][my rank]
    const double* slot_mine;                      // my slot[parity][0]
    unsigned long long* flag_peer[FEN_MAX_RANKS]; // peers' flag[CH_RED][my rank]
    const unsigned long long* flag_mine;          // my flag[CH_RED][0]
    int P, n, op;
    unsigned long long epoch, limit_ns;
    int* err;
};

// all-reduce of n <= kRedWidth doubles: every rank stores its values into everybody's slot, raises the
// flag, waits for all flags and reduces in rank order (so all ranks get bit-identical results).
__global__ void k_allreduce(RedArgs a) {
    const int r = threadIdx.x;
    if (r < a.P) {
        for (int q = 0; q < a.n; ++q) a.slot_peer[r][q] = a.vals[q];
        __threadfence_system();
        st_release_sys(a.flag_peer[r], a.epoch);
        spin_until(a.flag_mine + r, a.epoch, a.err, a.limit_ns);
    }
    __syncthreads();
    if (r < a.n) {
        double x = a.slot_mine[r];
        for (int q = 1; q < a.P; ++q) {
            const double y = a.slot_mine[(size_t)q * kRedWidth + r];
            x = a.op == 0 ? fmax(x, y) : x + y;
        }
        a.vals[r] = x;
    }
}

// ---- host side ------------------------------------------------------------------------------------
static unsigned long long* sig_ptr(Comm* m, int r, int ch, int slot) {
    return reinterpret_cast<unsigned long long*>(m->peer[r] + m->off_sig) + ch * kSigPerChannel + slot;
}
static int* err_ptr(Comm* m) { return m->h_err; }     // pinned, mapped: kernels write it, the host just reads it

static int need_comm(fen_ctx* c, Comm** out) {
    Comm* m = c->comm;
    if (!m || !m->connected)
        return set_error(FEN_ERR_COMM, "multi-GPU exchange: fen_gpu_comm_connect has not been called");
    *out = m;
    return FEN_OK;
}

int halo_exchange(fen_ctx* c, double* const* f, int n) {
    Comm* m;
    FEN_TRY(need_comm(c, &m));
    const Layout& L = c->L;
    const int P = m->P, rank = m->rank;
    const bool periodic = c->g.bc[4] == FEN_BC_PERIODIC && c->g.bc[5] == FEN_BC_PERIODIC;
    // neighbour on the lo (front) / hi (back) side; 2decomp wraps around when z is periodic
    int nb[2] = {rank - 1, rank + 1};
    if (nb[0] < 0) nb[0] = periodic ? P - 1 : -1;
    if (nb[1] >= P) nb[1] = periodic ? 0 : -1;
    const unsigned long long e = ++m->epoch[CH_HALO];
    const size_t par = (size_t)(e & 1);
    auto mail = [&](int r, int side, int comp) {
        return reinterpret_cast<double*>(m->peer[r] + m->off_mail) + ((par * 2 + side) * 3 + comp) * m->plane;
    };
    HaloArgs h;
    UnpackArgs u;
    memset(&h, 0, sizeof(h));
    memset(&u, 0, sizeof(u));
    h.n = u.n = n;
    h.plane2 = u.plane2 = m->plane / 2;
    for (int side = 0; side < 2; ++side) {
        h.has[side] = u.has[side] = nb[side] >= 0;
        for (int q = 0; q < n; ++q) {
            // my first interior plane goes to the lo neighbour's hi ghost, my last one to the hi neighbour's lo ghost
            h.src[side][q] = f[q] + (size_t)L.sz * (side == 0 ? 1 : L.nzl);
            h.dst[side][q] = nb[side] >= 0 ? mail(nb[side], 1 - side, q) : nullptr;
            u.src[side][q] = mail(rank, side, q);
            u.dst[side][q] = f[q] + (size_t)L.sz * (side == 0 ? 0 : L.nzl + 1);
        }
    }
    const int bx = (int)std::min<size_t>((h.plane2 + 255) / 256, 148 * 2);
    FEN_LAUNCH(c, "halo_push", k_halo_push<<<dim3(bx, 2, n), 256, 0, c->stream>>>(h));
    SigArgs s;
    memset(&s, 0, sizeof(s));
    s.epoch = e;
    s.limit_ns = spin_limit_ns();
    s.err = err_ptr(m);
    for (int side = 0; side < 2; ++side)
        if (nb[side] >= 0) {
            s.send_flag[s.nsend++] = sig_ptr(m, nb[side], CH_HALO, 1 - side);
            s.wait_flag[s.nwait++] = sig_ptr(m, rank, CH_HALO, side);
        }
    FEN_LAUNCH(c, "halo_sync", k_sigwait<<<1, 32, 0, c->stream>>>(s));
    FEN_LAUNCH(c, "halo_unpack", k_halo_unpack<<<dim3(bx, 2, n), 256, 0, c->stream>>>(u));
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

int comm_allreduce(fen_ctx* c, double* d_vals, int n, int op) {
    Comm* m;
    FEN_TRY(need_comm(c, &m));
    if (n > kRedWidth) return set_error(FEN_ERR_ARG, "allreduce of more than %d values", kRedWidth);
    const unsigned long long e = ++m->epoch[CH_RED];
    const size_t par = (size_t)(e & 1);
    RedArgs a;
    memset(&a, 0, sizeof(a));
    a.vals = d_vals;
    a.P = m->P; a.n = n; a.op = op; a.epoch = e; a.limit_ns = spin_limit_ns(); a.err = err_ptr(m);
    for (int r = 0; r < m->P; ++r) {
        a.slot_peer[r] = reinterpret_cast<double*>(m->peer[r] + m->off_red) + (par * FEN_MAX_RANKS + m->rank) * kRedWidth;
        a.flag_peer[r] = sig_ptr(m, r, CH_RED, m->rank);
    }
    a.slot_mine = reinterpret_cast<double*>(m->arena + m->off_red) + par * FEN_MAX_RANKS * kRedWidth;
    a.flag_mine = sig_ptr(m, m->rank, CH_RED, 0);
    FEN_LAUNCH(c, "allreduce", k_allreduce<<<1, 32, 0, c->stream>>>(a));
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

// the data of a transpose is pushed by the epilogue of the producing kernel (poisson.cu); this raises
// "my part is complete" on every peer and waits for theirs.  what: 1 = raise only, 2 = wait only (for the epoch the
// last raise of this channel used), 3 = both in one launch.
static int a2a_sync(fen_ctx* c, int ch, int what, cudaStream_t st, const char* name) {
    Comm* m;
    FEN_TRY(need_comm(c, &m));
    if (what & 1) ++m->epoch[ch];
    const unsigned long long e = m->epoch[ch];
    SigArgs s;
    memset(&s, 0, sizeof(s));
    s.epoch = e;
    s.limit_ns = spin_limit_ns();
    s.err = err_ptr(m);
    for (int r = 0; r < m->P; ++r) {
        if (r == m->rank) continue;
        if (what & 1) s.send_flag[s.nsend++] = sig_ptr(m, r, ch, m->rank);
        if (what & 2) s.wait_flag[s.nwait++] = sig_ptr(m, m->rank, ch, r);
    }
    FEN_LAUNCH(c, name, k_sigwait<<<1, 32, 0, st>>>(s));
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}
int comm_transpose_fwd(fen_ctx* c) { return a2a_sync(c, CH_A2A_FWD, 3, c->stream, "a2a_fwd_sync"); }
int comm_transpose_bwd(fen_ctx* c) { return a2a_sync(c, CH_A2A_BWD, 3, c->stream, "a2a_bwd_sync"); }
int comm_chunk_signal(fen_ctx* c, int q, cudaStream_t st) {
    if (q < 0 || q >= FEN_MAX_CHUNKS) return set_error(FEN_ERR_ARG, "transpose chunk %d", q);
    return a2a_sync(c, CH_A2A_CHUNK + q, 1, st, "a2a_bwd_sync");
}
int comm_chunk_wait(fen_ctx* c, int q, cudaStream_t st) {
    if (q < 0 || q >= FEN_MAX_CHUNKS) return set_error(FEN_ERR_ARG, "transpose chunk %d", q);
    return a2a_sync(c, CH_A2A_CHUNK + q, 2, st, "a2a_bwd_sync");
}

int comm_spectral(fen_ctx* c, double2** peerC, double2** peerCz) {
    Comm* m;
    FEN_TRY(need_comm(c, &m));
    for (int r = 0; r < m->P; ++r) {
        peerC[r] = reinterpret_cast<double2*>(m->peer[r] + m->off_C);
        peerCz[r] = reinterpret_cast<double2*>(m->peer[r] + m->off_Cz);
    }
    return FEN_OK;
}

int comm_check(fen_ctx* c) {
    Comm* m = c->comm;
    if (!m || !m->connected) return FEN_OK;
    // the stream has been synchronised by the caller
    const int e = *(volatile int*)m->h_err;
    if (e) return set_error(FEN_ERR_COMM, "a peer wait timed out after %llu ms (a rank is missing from a collective call)",
                            spin_limit_ns() / 1000000ull);
    return FEN_OK;
}

void comm_destroy(fen_ctx* c) {
    Comm* m = c->comm;
    if (!m) return;
    for (int r = 0; r < m->P; ++r)
        if (m->ipc_open[r] && m->peer[r]) cudaIpcCloseMemHandle(m->peer[r]);
    if (m->arena) cudaFree(m->arena);
    if (m->h_err) cudaFreeHost(m->h_err);
    delete m;
    c->comm = nullptr;
}

}  // namespace fen

using namespace fen;

extern "C" {

int fen_gpu_comm_handle_bytes(void) { return (int)sizeof(HandleBlob); }

int fen_gpu_comm_export(fen_ctx* c, void* out) {
    if (!c || !out) return set_error(FEN_ERR_ARG, "null argument");
    if (c->g.nranks < 2) return set_error(FEN_ERR_ARG, "comm_export on a single-rank context");
    if (c->g.nranks > FEN_MAX_RANKS) return set_error(FEN_ERR_UNSUPPORTED, "at most %d ranks", FEN_MAX_RANKS);
    FEN_CUDA(cudaSetDevice(c->device));
    if (c->g.ny % c->g.nranks)
        return set_error(FEN_ERR_UNSUPPORTED, "ny must be divisible by the number of ranks (z-pencil layout)");
    if (c->comm) comm_destroy(c);
    Comm* m = new Comm();
    c->comm = m;
    m->P = c->g.nranks;
    m->rank = c->g.rank;
    comm_layout(c, m);
    FEN_CUDA(cudaMalloc(&m->arena, m->arena_bytes));
    FEN_CUDA(cudaMemset(m->arena, 0, m->arena_bytes));
    FEN_CUDA(cudaDeviceSynchronize());
    FEN_CUDA(cudaHostAlloc(&m->h_err, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    *m->h_err = 0;
    HandleBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = kMagic;
    b.rank = m->rank; b.nranks = m->P; b.device = c->device;
    b.pid = (long long)getpid();
    b.arena_bytes = m->arena_bytes;
    b.raw = m->arena;
    gethostname(b.host, sizeof(b.host) - 1);
    FEN_CUDA(cudaIpcGetMemHandle(&b.ipc, m->arena));
    memcpy(out, &b, sizeof(b));
    return FEN_OK;
}

int fen_gpu_comm_connect(fen_ctx* c, const void* all) {
    if (!c || !all) return set_error(FEN_ERR_ARG, "null argument");
    Comm* m = c->comm;
    if (!m || !m->arena) return set_error(FEN_ERR_STATE, "comm_connect before comm_export");
    FEN_CUDA(cudaSetDevice(c->device));
    const HandleBlob* hb = static_cast<const HandleBlob*>(all);
    char host[64] = {0};
    gethostname(host, sizeof(host) - 1);
    for (int r = 0; r < m->P; ++r) {
        const HandleBlob& b = hb[r];
        if (b.magic != kMagic || b.rank != r || b.nranks != m->P)
            return set_error(FEN_ERR_COMM, "handle %d is not rank %d's export (gather them in rank order)", r, r);
        if (b.arena_bytes != m->arena_bytes)
            return set_error(FEN_ERR_COMM, "rank %d has a different grid (arena %llu vs %llu bytes)", r,
                             (unsigned long long)b.arena_bytes, (unsigned long long)m->arena_bytes);
        if (strncmp(b.host, host, sizeof(host)))
            return set_error(FEN_ERR_UNSUPPORTED, "rank %d is on host %s: peer memory needs all ranks on one NVLink box", r, b.host);
        if (r == m->rank) { m->peer[r] = m->arena; continue; }
        if (b.pid == (long long)getpid()) {        // same process (threads / one driver process)
            if (b.device != c->device) {
                int can = 0;
                FEN_CUDA(cudaDeviceCanAccessPeer(&can, c->device, b.device));
                if (!can) return set_error(FEN_ERR_COMM, "device %d cannot access device %d", c->device, b.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return set_error(FEN_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
                cudaGetLastError();
            }
            m->peer[r] = static_cast<char*>(b.raw);
        } else {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
                return set_error(FEN_ERR_COMM, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
            m->peer[r] = static_cast<char*>(p);
            m->ipc_open[r] = true;
        }
    }
    m->connected = true;
    return FEN_OK;
}

}  // extern "C"
