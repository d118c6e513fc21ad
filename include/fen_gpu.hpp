// fen_gpu.hpp -- header-only C++17 mirror of FEN's solver API over the C ABI of libfen_gpu.so (include/fen_gpu.h).
//
// The reference is compiled Fortran whose API is a set of derived types and module procedures: `type grid` with
// `setup` / `destroy` (src/grid.f90:22-62,67-200,269), `type scalar` with `allocate`, `update_ghost_nodes`, `max_value`,
// `integral`, `write`, `read` (src/scalar.f90:40-60), `type vector` (src/vector.f90), the module procedures
// `init_solver`, `advance_solution`, `save_state`, `load_state`, `destroy_solver` of solver_mod (src/solver.f90:34,75,160,
// 244,333), `set_timestep`, `print_solver_status` of navier_stokes_mod (src/navier_stokes.f90:623,734) and
// `init_poisson_solver` / `solve_poisson` (src/poisson.f90:51-57).  fortran/fen_gpu_mod.f90 is the bind(C) shim that
// keeps those names for Fortran drivers; this header keeps them for C++ hosts (the image has no Fortran compiler, so this
// is the compiled-language mirror that is built and run by the test-suite: examples/tgv_driver.cpp).  Same names, same
// argument meaning, errors as exceptions carrying fen_gpu_last_error().  Host arrays keep FEN's layout
// f(lo-gl:hi+gl, ...), x fastest; `push` / `pull` are the explicit transfer points that replace direct pokes into the
// reference's module-global arrays.  No torch, no CUDA types: plain pointers and sizes underneath.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#include "fen_gpu.h"

namespace fen {

struct error : std::runtime_error {
    int code;
    error(int c, const std::string& what_) : std::runtime_error(what_), code(c) {}
};
inline void check(int rc) {
    if (rc != FEN_OK) throw error(rc, fen_gpu_last_error());
}

// bc(6)%s of grid%setup: "Periodic", "Wall", "Inflow", "Outflow" (src/grid.f90:46)
inline int bc_code(const std::string& s) {
    if (s == "Periodic") return FEN_BC_PERIODIC;
    if (s == "Wall") return FEN_BC_WALL;
    if (s == "Inflow") return FEN_BC_INFLOW;
    if (s == "Outflow") return FEN_BC_OUTFLOW;
    throw error(FEN_ERR_ARG, "unknown boundary condition '" + s + "'");
}

// type grid (src/grid.f90:22-62)
class grid {
   public:
    int Nx = 0, Ny = 0, Nz = 0, ndim = 3, rank = 0, nranks = 1;
    double Lx = 0, Ly = 0, Lz = 0, delta = 0;
    std::array<int, 3> lo{{1, 1, 1}}, hi{{0, 0, 0}};
    fen_ctx* ctx = nullptr;

    grid() = default;
    grid(const grid&) = delete;
    grid& operator=(const grid&) = delete;
    ~grid() { if (ctx) fen_gpu_destroy(ctx); }

    // call comp_grid%setup(Nx, Ny, Nz, Lx, Ly, Lz, origin, prow, pcol, bc)   (grid.f90:67; prow is always 1: z slabs)
    void setup(int nx, int ny, int nz, double lx, double ly, double lz, int pcol = 1, int my_rank = 0,
               const std::vector<std::string>& bc = {}, int device = -1) {
        Nx = nx; Ny = ny; Nz = nz; Lx = lx; Ly = ly; Lz = lz;
        ndim = nz > 1 ? 3 : 2;
        rank = my_rank; nranks = pcol;
        delta = lx / static_cast<double>(static_cast<float>(nx));          // grid.f90:140: Lx/float(Nx), real*4
        fen_grid_desc d{};
        d.nx = nx; d.ny = ny; d.nz = nz; d.ndim = ndim; d.delta = delta;
        for (std::size_t q = 0; q < 6; ++q) d.bc[q] = q < bc.size() ? bc_code(bc[q]) : FEN_BC_PERIODIC;
        d.rank = my_rank; d.nranks = pcol; d.device = device;
        check(fen_gpu_create(&d, &ctx));
        check(fen_gpu_local_bounds(ctx, lo.data(), hi.data()));
    }
    void destroy() {                                                         // grid.f90:269
        if (ctx) check(fen_gpu_destroy(ctx));
        ctx = nullptr;
    }
    void synchronize() const { check(fen_gpu_synchronize(ctx)); }
    // cell-centre coordinates x(i), y(j), z(k), 1-based like the reference's (grid.f90:152-166, origin 0)
    double x(int i) const { return (i - 0.5) * delta; }
    double y(int j) const { return (j - 0.5) * delta; }
    double z(int k) const { return (k - 0.5) * delta; }
    int nloc(int dir) const { return hi[dir] - lo[dir] + 1; }
};

// type scalar (src/scalar.f90:40-60): host array f(lo-gl:hi+gl, ...) + its device twin
class scalar {
   public:
    grid* G = nullptr;
    int gl = 0, id = -1;
    bool owned = false;
    std::vector<double> f;
    std::array<int, 3> n{{0, 0, 0}};       // extents of f including ghosts

    scalar() = default;
    scalar(const scalar&) = delete;
    scalar& operator=(const scalar&) = delete;
    ~scalar() { if (owned && G && G->ctx && id >= 0) fen_gpu_scalar_destroy(G->ctx, id); }

    // call s%allocate(G, l)   (scalar.f90:63); field_id: one of the solver's own fields (FEN_VX ...) instead of a new one
    void allocate(grid& g, int l = 0, int loc = FEN_LOC_C, int field_id = -1) {
        G = &g; gl = l;
        n = {g.nloc(0) + 2 * l, g.nloc(1) + 2 * l, g.ndim == 3 ? g.nloc(2) + 2 * l : 1 + 2 * l};
        f.assign(static_cast<std::size_t>(n[0]) * n[1] * n[2], 0.0);
        if (field_id >= 0) { id = field_id; owned = false; }
        else { check(fen_gpu_scalar_allocate(g.ctx, l, loc, &id)); owned = true; }
    }
    // f(i, j, k) with the reference's global 1-based indices (ghosts: lo - gl .. hi + gl); 2-D: k = 1
    double& operator()(int i, int j, int k = 1) {
        const std::size_t a = static_cast<std::size_t>(i - G->lo[0] + gl), b = static_cast<std::size_t>(j - G->lo[1] + gl),
                          c = static_cast<std::size_t>(k - (G->ndim == 3 ? G->lo[2] : 1) + gl);
        return f[a + n[0] * (b + n[1] * c)];
    }
    void push() { check(fen_gpu_push(G->ctx, id, f.data(), gl)); }
    void pull() { check(fen_gpu_pull(G->ctx, id, f.data(), gl)); }
    void update_ghost_nodes() { check(fen_gpu_update_ghost_nodes(G->ctx, id, 1)); }      // scalar.f90:223
    void set_bc_type(int face, int type) { check(fen_gpu_set_bc_type(G->ctx, id, face, type)); }
    void set_bc(int face, double value) { check(fen_gpu_set_bc_plane(G->ctx, id, face, &value, 1)); }   // bc%<face> = value
    double max_value() { double v; check(fen_gpu_max_value(G->ctx, id, &v)); return v; }   // scalar.f90:179
    double integral() { double v; check(fen_gpu_integral(G->ctx, id, &v)); return v; }     // scalar.f90:201
    void write(const std::string& file) { check(fen_gpu_scalar_write(G->ctx, id, file.c_str())); }   // scalar.f90:428
    void read(const std::string& file) { check(fen_gpu_scalar_read(G->ctx, id, file.c_str())); }     // scalar.f90:400
};

// type vector (src/vector.f90): components x, y, z on the faces
class vector {
   public:
    grid* G = nullptr;
    scalar x, y, z;
    void allocate(grid& g, int l = 0, int first_id = -1) {                  // vector.f90:40
        G = &g;
        x.allocate(g, l, FEN_LOC_X, first_id < 0 ? -1 : first_id);
        y.allocate(g, l, FEN_LOC_Y, first_id < 0 ? -1 : first_id + 1);
        if (g.ndim == 3) z.allocate(g, l, FEN_LOC_Z, first_id < 0 ? -1 : first_id + 2);
    }
    void push() { x.push(); y.push(); if (G->ndim == 3) z.push(); }
    void pull() { x.pull(); y.pull(); if (G->ndim == 3) z.pull(); }
    void update_ghost_nodes() { check(fen_gpu_update_ghost_nodes(G->ctx, x.id, G->ndim)); }   // vector.f90:82
};

// solver_mod + navier_stokes_mod: module scalars as members, module procedures as methods
class solver {
   public:
    grid& G;
    double density = 1.0, viscosity = 1.0, CFL = 0.5;
    std::array<double, 3> g{{0.0, 0.0, 0.0}};
    bool constant_CFL = false;
    vector v, S;
    scalar p;

    explicit solver(grid& grid_) : G(grid_) {}
    ~solver() { if (ready_ && G.ctx) fen_gpu_destroy_solver(G.ctx); }

    void init_solver() {                                                     // solver.f90:34
        push_params();
        check(fen_gpu_init_solver(G.ctx));
        v.allocate(G, 1, FEN_VX);
        S.allocate(G, 0, FEN_SX);
        p.allocate(G, 1, FEN_LOC_C, FEN_P);
        ready_ = true;
    }
    double set_timestep(double U) {                                          // navier_stokes.f90:623
        push_params();
        double dt = 0.0;
        check(fen_gpu_set_timestep(G.ctx, U, &dt));
        return dt;
    }
    void advance_solution(int step, double& dt) {                            // solver.f90:75 -> navier_stokes.f90:50
        check(fen_gpu_navier_stokes_solver(G.ctx, step, &dt));
    }
    void status(double& maxdiv, double& maxCFL) { check(fen_gpu_get_status(G.ctx, &maxdiv, &maxCFL)); }
    std::string print_solver_status(int step, double time, double dt) {      // navier_stokes.f90:734
        char buf[256];
        check(fen_gpu_status_line(G.ctx, step, time, dt, buf, static_cast<int>(sizeof(buf))));
        return buf;
    }
    void save_state(const std::string& file) { check(fen_gpu_save_state(G.ctx, file.c_str())); }   // solver.f90:160
    void load_state(const std::string& file) { check(fen_gpu_load_state(G.ctx, file.c_str())); }   // solver.f90:244
    void save_fields(int step, const std::string& dir) { check(fen_gpu_save_fields(G.ctx, step, dir.c_str())); }
    std::string poisson_variant() const { return fen_gpu_poisson_variant(G.ctx); }
    void destroy_solver() {                                                  // solver.f90:333
        if (ready_) check(fen_gpu_destroy_solver(G.ctx));
        ready_ = false;
    }

   private:
    bool ready_ = false;
    void push_params() {
        fen_ns_params prm;
        check(fen_gpu_get_params(G.ctx, &prm));
        prm.density = density; prm.viscosity = viscosity; prm.CFL = CFL;
        prm.g[0] = g[0]; prm.g[1] = g[1]; prm.g[2] = g[2];
        prm.constant_CFL = constant_CFL ? 1 : 0;
        check(fen_gpu_set_params(G.ctx, &prm));
    }
};

// poisson_mod used on its own (test/small_test/poisson): init_Poisson_Solver(phi), solve_Poisson(phi)
class poisson_solver {
   public:
    explicit poisson_solver(scalar& phi) : G_(*phi.G) { check(fen_gpu_init_poisson_solver(G_.ctx)); }
    void solve(scalar& phi) { check(fen_gpu_solve_poisson(G_.ctx, phi.id)); }
    std::string variant() const { return fen_gpu_poisson_variant(G_.ctx); }
    void destroy() { check(fen_gpu_destroy_poisson_solver(G_.ctx)); }

   private:
    grid& G_;
};

}  // namespace fen
