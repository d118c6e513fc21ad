// io.cu -- raw field files in the reference's on-disk format, and the post-predictor host hook.
//
// FEN writes fields through 2decomp's MPI-IO: the GLOBAL interior array in natural Fortran order (x fastest),
// real(dp), no header (decomp_2d_write_one / decomp_2d_write_var; src/scalar.f90:400-455, src/solver.f90:103-329;
// postpro.py reads them with np.fromfile(...).reshape((Nx,Ny,Nz), order='F')).  With z slabs the part of rank r is
// one contiguous byte range of every field, so each rank preads / pwrites its own range of the shared file --
// no collective, no gather.
//   scalar%write / read          scalar.f90:428 / :400
//   save_state / load_state      solver.f90:160 / :244   order: p, v_x, v_y, dv_o_x, dv_o_y, [v_z, dv_o_z]
//                                                        two-phase build: + vof, rho, mu, p_o (:204-209, :288-297)
//   save_fields                  solver.f90:103          cell-centred velocities (face_to_center, fields.f90:210) and p
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstring>

#include "fen_internal.cuh"

namespace fen {

struct IoBuf {     // view of the context's pinned host staging for one slab interior (allocated once, fen_ctx::io_host)
    double* h = nullptr;
    size_t n = 0;
};

static size_t slab_elems(fen_ctx* c) { return (size_t)c->L.nx * c->L.ny * c->L.nzl; }
static long long slab_offset(fen_ctx* c, int nfield_before) {
    const long long global = (long long)c->g.nx * c->g.ny * c->g.nz;
    return ((long long)nfield_before * global + (long long)c->k0 * c->g.nx * c->g.ny) * (long long)sizeof(double);
}

static int xfer_all(int fd, bool write, char* buf, size_t bytes, long long off, const char* path) {
    while (bytes > 0) {
        const ssize_t r = write ? pwrite(fd, buf, bytes, off) : pread(fd, buf, bytes, off);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0)
            return set_error(FEN_ERR_ARG, "%s %s failed at offset %lld: %s", write ? "writing" : "reading", path, off,
                             r == 0 ? "unexpected end of file" : strerror(errno));
        buf += r; bytes -= (size_t)r; off += r;
    }
    return FEN_OK;
}

// one field <-> its byte range in an open file (field number `slot` of the file)
static int field_to_file(fen_ctx* c, int fd, const char* path, int id, int slot, IoBuf& b) {
    FEN_TRY(fen_gpu_pull(c, id, b.h, 0));                    // interior only, synchronises
    return xfer_all(fd, true, reinterpret_cast<char*>(b.h), b.n * sizeof(double), slab_offset(c, slot), path);
}
static int file_to_field(fen_ctx* c, int fd, const char* path, int id, int slot, IoBuf& b) {
    FEN_TRY(xfer_all(fd, false, reinterpret_cast<char*>(b.h), b.n * sizeof(double), slab_offset(c, slot), path));
    FEN_TRY(fen_gpu_push(c, id, b.h, 0));
    FEN_CUDA(cudaStreamSynchronize(c->stream));             // the staging buffer is reused for the next field
    return FEN_OK;
}

static int io_begin(fen_ctx* c, const char* path, bool write, int nfields, int* fd, IoBuf& b) {
    if (!c || !path) return set_error(FEN_ERR_ARG, "null argument");
    FEN_CUDA(cudaSetDevice(c->device));
    b.n = slab_elems(c);
    if (c->io_host_n < b.n) {
        // first I/O call of this context (or a larger slab): the only allocation the I/O path ever makes.  Like every
        // allocation it synchronises the device, so on several ranks call the first save / load between steps, after a
        // caller-side barrier -- as the reference's collective MPI-IO calls are (solver.f90:160, :244)
        if (c->io_host) cudaFreeHost(c->io_host);
        c->io_host = nullptr; c->io_host_n = 0;
        FEN_CUDA(cudaMallocHost(&c->io_host, b.n * sizeof(double)));
        c->io_host_n = b.n;
    }
    b.h = c->io_host;
    *fd = write ? open(path, O_CREAT | O_WRONLY, 0644) : open(path, O_RDONLY);
    if (*fd < 0) return set_error(FEN_ERR_ARG, "cannot open %s: %s", path, strerror(errno));
    if (write && c->g.rank == 0) {
        // exact final size (the reference truncates to guarantee overwriting, solver.f90:196); ranges already
        // written by other ranks are inside the new size and stay intact
        const long long total = (long long)nfields * c->g.nx * c->g.ny * c->g.nz * (long long)sizeof(double);
        if (ftruncate(*fd, total) != 0) return set_error(FEN_ERR_ARG, "cannot size %s: %s", path, strerror(errno));
    }
    return FEN_OK;
}

__global__ void __launch_bounds__(256) k_face_to_center(Layout L, const double* f, double* o, long long back) {
    // fields.f90:210-245: sc = 0.5 (sf + sf shifted by one towards the low side)
    const int j = blockIdx.y + 1, k = blockIdx.z + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x + 1; i <= L.nx; i += gridDim.x * blockDim.x) {
        const long long cidx = L.idx(i, j, k);
        o[cidx] = 0.5 * (f[cidx] + f[cidx - back]);
    }
}

}  // namespace fen

using namespace fen;

extern "C" {

int fen_gpu_scalar_write(fen_ctx* c, int field, const char* filename) {
    int fd = -1;
    IoBuf b;
    int r = io_begin(c, filename, true, 1, &fd, b);
    if (r == FEN_OK) r = field_to_file(c, fd, filename, field, 0, b);
    if (fd >= 0) close(fd);
    return r;
}

int fen_gpu_scalar_read(fen_ctx* c, int field, const char* filename) {
    int fd = -1;
    IoBuf b;
    int r = io_begin(c, filename, false, 1, &fd, b);
    if (r == FEN_OK) r = file_to_field(c, fd, filename, field, 0, b);
    if (fd >= 0) close(fd);
    return r;
}

static int state_fields(fen_ctx* c, int* ids) {
    int n = 0;
    ids[n++] = FEN_P; ids[n++] = FEN_VX; ids[n++] = FEN_VY; ids[n++] = FEN_DVOX; ids[n++] = FEN_DVOY;
    if (c->g.ndim == 3) { ids[n++] = FEN_VZ; ids[n++] = FEN_DVOZ; }
    if (mf_active(c)) { ids[n++] = FEN_VOF; ids[n++] = FEN_RHO; ids[n++] = FEN_MU; ids[n++] = FEN_PO; }   // solver.f90:204-209
    return n;
}

int fen_gpu_save_state(fen_ctx* c, const char* filename) {
    if (!c || !c->solver_init) return set_error(FEN_ERR_STATE, "init_solver has not been called");
    int ids[12];
    const int n = state_fields(c, ids);
    int fd = -1;
    IoBuf b;
    int r = io_begin(c, filename, true, n, &fd, b);
    for (int q = 0; q < n && r == FEN_OK; ++q) r = field_to_file(c, fd, filename, ids[q], q, b);
    if (fd >= 0) close(fd);
    return r;
}

int fen_gpu_load_state(fen_ctx* c, const char* filename) {
    if (!c || !c->solver_init) return set_error(FEN_ERR_STATE, "init_solver has not been called");
    int ids[12];
    const int n = state_fields(c, ids);
    int fd = -1;
    IoBuf b;
    int r = io_begin(c, filename, false, n, &fd, b);
    if (r == FEN_OK) {
        struct stat st;
        const long long need = (long long)n * c->g.nx * c->g.ny * c->g.nz * (long long)sizeof(double);
        if (fstat(fd, &st) != 0 || (long long)st.st_size != need)
            r = set_error(FEN_ERR_ARG, "%s is not a state file of this grid (%lld bytes expected)", filename, need);
    }
    for (int q = 0; q < n && r == FEN_OK; ++q) r = file_to_field(c, fd, filename, ids[q], q, b);
    if (fd >= 0) close(fd);
    if (r != FEN_OK) return r;
    FEN_TRY(ghost_update(c, FEN_P, 1));                       // solver.f90:283-284
    FEN_TRY(ghost_update(c, FEN_VX, c->g.ndim));
    if (mf_active(c)) {                                       // solver.f90:293-296
        FEN_TRY(ghost_update(c, FEN_VOF, 1));
        FEN_TRY(ghost_update(c, FEN_RHO, 2));
        FEN_TRY(ghost_update(c, FEN_PO, 1));
    }
    return FEN_OK;
}

// save_fields(step): <dir>/vx_<step7>.raw, vy_, [vz_], p_  (solver.f90:103-156)
int fen_gpu_save_fields(fen_ctx* c, int step, const char* dir) {
    if (!c || !c->solver_init) return set_error(FEN_ERR_STATE, "init_solver has not been called");
    if (!dir) dir = "data";
    if (c->io_tmp < 0) FEN_TRY(fen_gpu_scalar_allocate(c, 1, FEN_LOC_C, &c->io_tmp));   // kept for the next call
    const int tmp = c->io_tmp;
    const char* names[3] = {"vx", "vy", "vz"};
    const Layout& L = c->L;
    const long long back[3] = {1, L.sy, L.sz};
    char path[1024];
    int r = FEN_OK;
    for (int m = 0; m < c->g.ndim && r == FEN_OK; ++m) {
        Field *v, *t;
        r = field_check(c, FEN_VX + m, &v);
        if (r == FEN_OK) r = field_check(c, tmp, &t);
        if (r != FEN_OK) break;
        dim3 grid((L.nx + 255) / 256, L.ny, L.nzl), block(256);
        FEN_LAUNCH(c, "face_to_center", k_face_to_center<<<grid, block, 0, c->stream>>>(L, v->d, t->d, back[m]));
        snprintf(path, sizeof(path), "%s/%s_%07d.raw", dir, names[m], step);
        r = fen_gpu_scalar_write(c, tmp, path);
    }
    if (r == FEN_OK) {
        snprintf(path, sizeof(path), "%s/p_%07d.raw", dir, step);
        r = fen_gpu_scalar_write(c, FEN_P, path);
    }
    if (r == FEN_OK && mf_active(c)) {                        // solver.f90:145-148
        snprintf(path, sizeof(path), "%s/vof_%07d.raw", dir, step);
        r = fen_gpu_scalar_write(c, FEN_VOF, path);
    }
    return r;
}

int fen_gpu_set_forcing_hook(fen_ctx* c, fen_forcing_fn fn, void* user) {
    if (!c) return set_error(FEN_ERR_ARG, "null context");
    c->forcing = fn;
    c->forcing_user = user;
    return FEN_OK;
}

}  // extern "C"
