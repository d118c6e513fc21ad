// fen_internal.cuh -- internal types of libfen_gpu.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <tuple>
#include <map>
#include <string>
#include <vector>

#include "../../include/fen_gpu.h"

#define FEN_MAX_RANKS 16
#define FEN_MAX_CHUNKS 8      // pieces of a chunked slab transpose (comm.cu: one flag channel each)

namespace fen {

// ---------------------------------------------------------------------------------------------
// Device layout of every real field (HBM): the reference's ghosted x-pencil array
// f(0:Nx+1, 0:Ny+1, lo3-1:hi3+1) (src/scalar.f90:79-81) with the x rows padded so that the first
// INTERIOR element of every row sits on a 128-byte boundary:
//     element (i, j, k), i in [0, nx+1], j in [0, ny+1], k in [0, nzl+1]  (k local to the slab)
//     lives at  (xoff - 1 + i) + px * (j + (ny + 2) * k),   xoff = 16, px = roundup(xoff+nx+1, 16)
// Fields the reference allocates without ghosts (gl = 0) use the same layout; their ghost cells
// are simply never read.
// ---------------------------------------------------------------------------------------------
struct Layout {
    int nx, ny, nzl;
    int px, xoff;
    long long sy, sz;
    size_t elems;
    __host__ __device__ __forceinline__ long long idx(int i, int j, int k) const {
        return (long long)(xoff - 1 + i) + sy * j + sz * k;
    }
};

enum BcMode { BC_ZERO = 0, BC_UNIFORM = 1, BC_PLANE = 2 };

struct Field {
    double* d = nullptr;
    int gl = 0;
    int loc = FEN_LOC_C;
    bool exists = false;       // id handed out
    int bc_type[6] = {0, 0, 0, 0, 0, 0};
    int bc_mode[6] = {0, 0, 0, 0, 0, 0};
    double bc_value[6] = {0, 0, 0, 0, 0, 0};
    double* bc_plane[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pull_event = nullptr;   // fen_gpu_pull_async: the host copy of this field is still in flight until it fires
    // chunked asynchronous pull (FEN_COPY_CHUNKS > 1): the copy went out in pull_chunks pieces, each with its own event, so
    // that a later push of the same host array can follow it piece by piece (context.cu: copy_field)
    int pull_chunks = 0, pull_buf = 0;
    const double* pull_host = nullptr;
    size_t pull_n = 0;
};

struct ProfEntry {
    int name_id;
    cudaEvent_t e0, e1;
};

// cache key of a field's TMA descriptor: buffer and box shape
struct TmapKey {
    const double* base;
    int bx, by;
    bool operator<(const TmapKey& o) const { return std::tie(base, bx, by) < std::tie(o.base, o.bx, o.by); }
};

// one captured navier_stokes_solver step (context.cu): replayed while the step's inputs keep the same signature
struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;    // kernels inside, for fen_gpu_launch_count
    bool net_swap = false;     // the step leaves v in the other ping-pong buffer
    int seen = 0;
};

// Kernel attributes (the dynamic shared-memory limit above 48 KB) belong to the device a kernel is loaded on, not to the
// process: a caller may drive several contexts -- several GPUs, or several ranks on one GPU -- from one process, each
// from its own thread.  The one-time set-up of a launcher is therefore tracked per device AND serialised: the thread
// that finds the bit clear holds the lock until it leaves the launcher's scope, so no other thread can launch the kernel
// before its attributes are set (a bare check-and-set let a second rank launch with 128 KB of dynamic shared memory a
// moment before the first had raised the limit: "invalid argument", seen with 8 ranks as threads).
// Usage:  FEN_ONCE_PER_DEVICE(c) { FEN_CUDA(cudaFuncSetAttribute(...)); }
struct OnceState {
    std::mutex m;
    std::atomic<unsigned long long> done{0};
};
class OnceLock {
   public:
    OnceLock(OnceState& st, int device) : st_(st), bit_(1ull << (device & 63)) {
        if (st_.done.load(std::memory_order_acquire) & bit_) return;
        st_.m.lock();
        if (st_.done.load(std::memory_order_acquire) & bit_) { st_.m.unlock(); return; }
        first_ = true;
    }
    ~OnceLock() {
        if (first_) {
            st_.done.fetch_or(bit_, std::memory_order_release);
            st_.m.unlock();
        }
    }
    bool first() const { return first_; }
    OnceLock(const OnceLock&) = delete;
    OnceLock& operator=(const OnceLock&) = delete;

   private:
    OnceState& st_;
    unsigned long long bit_;
    bool first_ = false;
};
#define FEN_ONCE_PER_DEVICE(ctx)                            \
    static fen::OnceState once_state__;                     \
    fen::OnceLock once_lock__(once_state__, (ctx)->device); \
    if (once_lock__.first())

struct Poisson;   // poisson.cu
struct Comm;      // comm.cu

// state of volume_of_fluid_mod / multiphase_mod (multiphase.cu); allocated by fen_gpu_allocate_vof_fields
struct Multiphase {
    fen_mf_params prm{};
    bool vof_fields = false;   // allocate_vof_fields has run
    bool ns = false;           // init_solver_mf has run: the step takes the -DMF branches
};

}  // namespace fen

struct fen_ctx {
    fen_grid_desc g{};
    int device = 0;
    cudaStream_t stream = nullptr;
    fen::Layout L{};
    int k0 = 0;                 // global index of local plane k = 1 minus 1 (lo(3) - 1)
    std::vector<fen::Field> fields;

    // navier_stokes_mod state
    bool solver_init = false;
    fen_ns_params prm{};
    double rho_uniform = 1.0, mu_uniform = 1.0;   // value of rho%f / mu%f while they are uniform
    bool uniform_props = true;                    // rho and mu are uniform fields (hazard H11)
    bool has_source = false;                      // S was written by the caller
    double* vnew[3] = {nullptr, nullptr, nullptr};   // predictor output (ping-pong with v)
    double maxdiv = 0.0, maxCFL = 0.0;
    double last_dt = 0.0;
    double* stage = nullptr;     // contiguous staging buffer of push / pull (context.cu: copy_field)
    // fen_gpu_pull_async: device-to-host copies on their own stream, four staging buffers in rotation, so that the
    // download of one field overlaps the repitch of the next and -- PCIe being full duplex -- later uploads
    cudaStream_t d2h = nullptr;
    double* stage_out[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_ready[4] = {nullptr, nullptr, nullptr, nullptr};   // repitch into stage_out[b] done (compute stream)
    cudaEvent_t ev_free[4] = {nullptr, nullptr, nullptr, nullptr};    // host copy out of stage_out[b] done (d2h stream)
    bool ev_free_set[4] = {false, false, false, false};
    cudaEvent_t ev_chunk[4][8] = {};   // per staging buffer: piece q of its host copy is out (d2h stream)
    int out_next = 0;
    // io.cu: pinned staging of one slab interior and the cell-centred temporary of save_fields, allocated on first use
    // and kept -- so that the I/O entry points do not allocate (a device-wide synchronisation) on every call
    double* io_host = nullptr;
    size_t io_host_n = 0;
    int io_tmp = -1;
    double* d_red = nullptr;     // device scratch for reductions (partials + results)
    double* h_red = nullptr;     // pinned host mirror of the results
    int red_blocks = 0;

    std::map<unsigned long long, fen::StepGraph> step_graphs;   // CUDA graphs of the step, keyed by its signature
    bool graphs_off = false;            // capture failed once, or FEN_NO_GRAPH: always enqueue eagerly
    fen_forcing_fn forcing = nullptr;   // host hook after the predictor (io.cu)
    void* forcing_user = nullptr;

    fen::Poisson* ps = nullptr;
    fen::Comm* comm = nullptr;
    fen::Multiphase* mf = nullptr;
    std::map<fen::TmapKey, CUtensorMap> tmaps;      // TMA descriptors of the field buffers (tma.cu)

    // measurement
    long long launches = 0;
    bool profiling = false;
    std::vector<std::string> prof_names;
    std::vector<fen::ProfEntry> prof_entries;
};

namespace fen {

int set_error(int code, const char* fmt, ...);
#define FEN_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fen::set_error(FEN_ERR_CUDA, "%s failed: %s (%s:%d)", #call,              \
                                  cudaGetErrorString(e__), __FILE__, __LINE__);              \
    } while (0)
#define FEN_TRY(call)                 \
    do {                              \
        int r__ = (call);             \
        if (r__ != FEN_OK) return r__; \
    } while (0)

// RAII-less launch bracket: counts launches and, when profiling, brackets the launch with events.
int prof_begin(fen_ctx* c, const char* name);
void prof_end(fen_ctx* c, int token);
#define FEN_LAUNCH(ctx, name, ...)               \
    do {                                         \
        int tok__ = fen::prof_begin(ctx, name);  \
        __VA_ARGS__;                             \
        fen::prof_end(ctx, tok__);               \
    } while (0)

int field_check(fen_ctx* c, int id, Field** out, bool alloc = true);
int field_alloc(fen_ctx* c, Field& f);
void init_field(fen_ctx* c, int id, int gl, int loc);     // scalar%allocate defaults (scalar.f90:63-133)
void free_field(Field& f);
void step_graphs_clear(fen_ctx* c);                        // drops the captured step graphs (they bake buffer addresses)
int fetch_red(fen_ctx* c, int n);                          // async copy of d_red[0..n) to h_red

// tma.cu
int field_tmap(fen_ctx* c, const double* base, int box_x, int box_y, CUtensorMap** out);
// ghost.cu
int ghost_update(fen_ctx* c, int field, int ncomp, bool x_done = false);   // x_done: producer wrote periodic x ghosts
int ghost_update_list(fen_ctx* c, const int* ids, int n, bool x_done = false);   // up to 8 fields, one launch per direction
// comm.cu
int halo_exchange(fen_ctx* c, double* const* f, int n);
int comm_allreduce(fen_ctx* c, double* d_vals, int n, int op /*0 max, 1 sum*/);
void comm_destroy(fen_ctx* c);
int comm_transpose_fwd(fen_ctx* c);   // completes the y-slab -> z-pencil transpose pushed by the y FFT
int comm_transpose_bwd(fen_ctx* c);   // completes the z-pencil -> y-slab transpose pushed by the z stage
int comm_chunk_signal(fen_ctx* c, int q, cudaStream_t st);   // chunk q of a chunked z -> y transpose: my stores are out
int comm_chunk_wait(fen_ctx* c, int q, cudaStream_t st);     // ... and everybody else's have arrived
int comm_spectral(fen_ctx* c, double2** peerC, double2** peerCz);   // mapped spectral arrays of all ranks
int comm_check(fen_ctx* c);           // FEN_ERR_COMM if a peer wait timed out (call after a stream sync)
int spectral_pitch(int nx);           // complex row pitch of the half-spectrum arrays
int spectral_pitch_grid(const fen_grid_desc& g);   // ... of this grid's Poisson variant (full width for DCT in x)
// stencil.cu
int ns_predict(fen_ctx* c, double dt);
int ns_poisson_rhs(fen_ctx* c, double dt);
int ns_correct(fen_ctx* c, double dt, bool* checks_done = nullptr);   // checks_done: fused checks ran
int ns_checks_launch(fen_ctx* c, double dt);    // leaves (maxdiv, maxvel) in d_red[0..1]
int op_gradient(fen_ctx* c, int s, int vx);
int op_divergence(fen_ctx* c, int vx, int s);
int op_laplacian(fen_ctx* c, int vx, int ox);
int op_center_to_face(fen_ctx* c, int s, int vx);
int op_laplacian_scalar(fen_ctx* c, int s, int o);
int op_face_to_center(fen_ctx* c, int sf, int sc, int dir);
int op_curl(fen_ctx* c, int vx, int ox);
int op_explicit_terms(fen_ctx* c, int rhs_x, bool advection_only);
int ensure_red(fen_ctx* c);
int reduce_field(fen_ctx* c, const double* f, int op, double* d_out);   // op 0 max, 1 sum
int field_is_uniform(fen_ctx* c, const double* f, bool* uniform, double* value);
// poisson.cu
int poisson_init(fen_ctx* c);
// fuse_rhs: the x pass computes rhs = div(v) rho/dt from the velocity field itself instead of reading f
int poisson_solve(fen_ctx* c, double* f, bool fuse_rhs = false, double dt = 0.0);
bool poisson_can_fuse_rhs(fen_ctx* c);
void poisson_destroy(fen_ctx* c);
const char* poisson_variant(fen_ctx* c);
// multiphase.cu
inline bool mf_active(const fen_ctx* c) { return c->mf && c->mf->ns; }
int mf_step_front(fen_ctx* c, double dt);      // advect_interface, material properties, p_hat (navier_stokes.f90:80-96)
int mf_predict(fen_ctx* c, double dt);         // predicted_velocity_field with the MF terms (:140-213, :405-501)
int mf_poisson_rhs(fen_ctx* c, double dt);     // phi = div(v) rhomin/dt (:111-113)
int mf_correct(fen_ctx* c, double dt);         // correct_velocity_field + update_pressure, MF branches (:526-531, :553-564)
int mf_set_timestep(fen_ctx* c, double U, double* dt);
void mf_destroy(fen_ctx* c);

}  // namespace fen
