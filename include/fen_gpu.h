/*
 * fen_gpu.h -- flat C ABI of libfen_gpu.so: the B200 (sm_100a) drop-in for FEN's fractional-step
 * Navier-Stokes hot path (predictor stencils -> FFT/tridiagonal Poisson -> projection, with the
 * ghost-cell exchange and slab transposes).
 *
 * FEN has no FFI today; its boundary is a set of Fortran module procedures and module-global
 * fields.  Every entry point below names the reference procedure (file:line under
 * /root/reference) it replaces; fortran/fen_gpu_mod.f90 holds the bind(C) interfaces and the thin
 * Fortran wrappers with the reference's own names, and INTEGRATION.md shows how a maintainer
 * wires them in.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; fen_gpu_last_error() gives the text.
 *    The reference prints errors to stderr and continues (src/IO.f90:12-30) except for an
 *    unsupported Poisson BC combination, which stops (src/poisson.f90:91-95): here that case is
 *    FEN_ERR_UNSUPPORTED from fen_gpu_init_solver / fen_gpu_init_poisson_solver.
 *  - host arrays are Fortran-ordered real(dp), x fastest, exactly the reference's
 *    f(lo(1)-gl:hi(1)+gl, lo(2)-gl:hi(2)+gl, lo(3)-gl:hi(3)+gl) (src/scalar.f90:79-81) for the
 *    calling rank's x-pencil; pass c_loc(f) and gl.
 *  - decomposition: FEN's own 2decomp layout with (prow, pcol) = (1, nranks): x and y whole,
 *    z split in nranks equal slabs (src/grid.f90:125,168-173).  One context per GPU / rank.
 *  - a context is not thread-safe (neither are the reference's module globals); calls are
 *    enqueued on the context's CUDA stream in call order.  Multi-rank calls are collective: all
 *    ranks make them in the same order (as with the reference's MPI calls).
 *  - there is no CPU fallback: without a CUDA device fen_gpu_create fails.
 */
#ifndef FEN_GPU_H
#define FEN_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fen_ctx fen_ctx;

enum fen_status {
    FEN_OK = 0,
    FEN_ERR_ARG = 1,          /* bad argument / unknown field */
    FEN_ERR_CUDA = 2,         /* CUDA runtime error */
    FEN_ERR_UNSUPPORTED = 3,  /* outside the hot-path scope (BC combo, transform length with a prime factor > 61 ...) */
    FEN_ERR_STATE = 4,        /* call order (e.g. step before init_solver) */
    FEN_ERR_COMM = 5          /* multi-GPU exchange set-up */
};

/* Grid boundary strings of src/grid.f90:46 / navier_stokes.f90:780-1017 */
enum fen_grid_bc { FEN_BC_PERIODIC = 0, FEN_BC_WALL = 1, FEN_BC_INFLOW = 2, FEN_BC_OUTFLOW = 3 };

/* Field BC type codes of src/scalar.f90:17-22 */
enum fen_bc_type { FEN_HALO = -1, FEN_PERIODIC = 0, FEN_DIRICHLET = 1, FEN_NEUMANN = 2 };

/* Faces in the reference's order: left, right (x), bottom, top (y), front, back (z) */
enum fen_face { FEN_LEFT = 0, FEN_RIGHT = 1, FEN_BOTTOM = 2, FEN_TOP = 3, FEN_FRONT = 4, FEN_BACK = 5 };

/* Location tag `c` of type scalar (src/scalar.f90:46) */
enum fen_loc { FEN_LOC_C = 0, FEN_LOC_X = 1, FEN_LOC_Y = 2, FEN_LOC_Z = 3 };

/* Module-global fields of navier_stokes_mod (src/navier_stokes.f90:36-38).  A vector component
 * is base + {0,1,2}.  Ids >= FEN_FIELD_USER come from fen_gpu_scalar_allocate. */
enum fen_field {
    FEN_P = 0, FEN_PHI = 1, FEN_RHO = 2, FEN_MU = 3,
    FEN_VX = 4, FEN_VY = 5, FEN_VZ = 6,
    FEN_DVX = 7, FEN_DVY = 8, FEN_DVZ = 9,          /* dv      */
    FEN_DVOX = 10, FEN_DVOY = 11, FEN_DVOZ = 12,    /* dv_o    */
    FEN_GPX = 13, FEN_GPY = 14, FEN_GPZ = 15,       /* grad_p  */
    FEN_SX = 16, FEN_SY = 17, FEN_SZ = 18,          /* S       */
    /* two-phase build (-DMF): module fields of volume_of_fluid_mod (volume_of_fluid.f90:21-22) and of
     * multiphase_mod (multiphase.f90:26); they exist after fen_gpu_allocate_vof_fields / fen_gpu_init_solver_mf */
    FEN_VOF = 19, FEN_H = 20, FEN_D = 21, FEN_CURV = 22,
    FEN_NORMX = 23, FEN_NORMY = 24, FEN_NORMZ = 25, /* norm    */
    FEN_LX = 26, FEN_LY = 27, FEN_LZ = 28,          /* l       */
    FEN_PHAT = 29, FEN_PO = 30,                     /* p_hat, p_o */
    FEN_VOF1 = 31,                                  /* advect_vof's temporary vof1 (volume_of_fluid.f90:445) */
    FEN_FIELD_USER = 32
};

/* type grid + grid%setup (src/grid.f90:22-62, 67-200), reduced to what the path needs. */
typedef struct fen_grid_desc {
    int nx, ny, nz;       /* global cells */
    int ndim;             /* cpp macro DIM: 2 (nz must be 1) or 3 */
    double delta;         /* grid%delta = Lx/float(Nx), computed by the caller (grid.f90:140) */
    int bc[6];            /* fen_grid_bc per face; entries 4,5 ignored when ndim == 2 */
    int rank, nranks;     /* slab owner / number of z slabs (pcol); prow is always 1 */
    int device;           /* CUDA device ordinal, -1 = current device */
} fen_grid_desc;

const char* fen_gpu_last_error(void);
int fen_gpu_version(void);

/* ---- grid / context: grid%setup (grid.f90:67), grid%destroy (:269) ------------------------ */
int fen_gpu_create(const fen_grid_desc* desc, fen_ctx** out);
int fen_gpu_destroy(fen_ctx* ctx);
int fen_gpu_synchronize(fen_ctx* ctx);
/* local x-pencil bounds lo(3), hi(3) (1-based, grid.f90:168-169) */
int fen_gpu_local_bounds(fen_ctx* ctx, int lo[3], int hi[3]);

/* ---- multi-GPU set-up (replaces decomp_2d_init, grid.f90:125) ------------------------------
 * Each rank exports fen_gpu_comm_handle_bytes() bytes with fen_gpu_comm_export, the host side
 * all-gathers them (rank order) by any means (MPI, torch.distributed, files) and gives every rank
 * the concatenation through fen_gpu_comm_connect.  After that halos and transposes move over
 * NVLink peer memory inside this library's own kernels. */
int fen_gpu_comm_handle_bytes(void);
int fen_gpu_comm_export(fen_ctx* ctx, void* handle_out);
int fen_gpu_comm_connect(fen_ctx* ctx, const void* all_handles);

/* ---- scalar / vector containers (src/scalar.f90, src/vector.f90) -------------------------- */
/* scalar%allocate (scalar.f90:63): returns a new field id in *field */
int fen_gpu_scalar_allocate(fen_ctx* ctx, int gl, int loc, int* field);
int fen_gpu_scalar_destroy(fen_ctx* ctx, int field);
/* host f(...) -> device, device -> host; gl = ghost layers of the HOST array (0 or 1) */
int fen_gpu_push(fen_ctx* ctx, int field, const double* host, int gl);
int fen_gpu_pull(fen_ctx* ctx, int field, double* host, int gl);
/* pull without blocking: the copy runs on its own stream (PCIe is full duplex, so it overlaps later pushes); the host
 * array is valid after fen_gpu_pull_wait or fen_gpu_synchronize.  A later push of the same field waits for it on the
 * device side. */
int fen_gpu_pull_async(fen_ctx* ctx, int field, double* host, int gl);
int fen_gpu_pull_wait(fen_ctx* ctx);
/* self%f = val (scalar%setToValue, scalar.f90:168) */
int fen_gpu_set_to_value(fen_ctx* ctx, int field, double val);
/* bc%type_<face> (scalar.f90:24-37) */
int fen_gpu_set_bc_type(fen_ctx* ctx, int field, int face, int type);
int fen_gpu_get_bc_type(fen_ctx* ctx, int field, int face, int* type);
/* bc%<face>(:,:) value plane incl. ghosts, Fortran order (scalar.f90:89-96); NULL = all zero.
 * `uniform` != 0: plane[0] is broadcast (e.g. v%x%bc%top = U, lid_driven.f90:59). */
int fen_gpu_set_bc_plane(fen_ctx* ctx, int field, int face, const double* plane, int uniform);
/* scalar%update_ghost_nodes (scalar.f90:223) / vector%update_ghost_nodes (vector.f90:82):
 * halo exchange + physical BCs in the reference's order.  ncomp = 1 (scalar) or ndim (vector,
 * field = the x component). */
int fen_gpu_update_ghost_nodes(fen_ctx* ctx, int field, int ncomp);
/* update_halos(f, G, l) (src/halo.f90:12): z-neighbour exchange only (incl. periodic wrap when
 * nranks > 1), level 1 */
int fen_gpu_update_halos(fen_ctx* ctx, int field);
/* scalar%max_value (scalar.f90:179) and scalar%integral (:201), reduced over all ranks */
int fen_gpu_max_value(fen_ctx* ctx, int field, double* out);
int fen_gpu_integral(fen_ctx* ctx, int field, double* out);

/* ---- fields_mod (src/fields.f90) ----------------------------------------------------------- */
int fen_gpu_gradient(fen_ctx* ctx, int scalar_in, int vector_out_x);        /* fields.f90:31  */
int fen_gpu_divergence(fen_ctx* ctx, int vector_in_x, int scalar_out);      /* fields.f90:120 */
int fen_gpu_laplacian(fen_ctx* ctx, int vector_in_x, int vector_out_x);     /* fields.f90:298 */
int fen_gpu_center_to_face(fen_ctx* ctx, int scalar_in, int vector_out_x);  /* fields.f90:175 */
/* the other operators of the module's generic interfaces, used by the callers either side of the step: laplacian(s, lap_s)
 * (the reference's fields test, test/small_test/fields/methods.f90:326), face_to_center (save_fields, solver.f90:131-140;
 * dir 0 / 1 / 2 = the reference's 'x' / 'y' / 'z'), curl (lid_driven.f90:86; 2-D: the z component goes to the x
 * component of the output, fields.f90:381-384).  Inputs need ghost nodes; an output must not be an input. */
int fen_gpu_laplacian_scalar(fen_ctx* ctx, int scalar_in, int scalar_out);             /* fields.f90:256 */
int fen_gpu_face_to_center(fen_ctx* ctx, int scalar_face, int scalar_center, int dir);  /* fields.f90:210 */
int fen_gpu_curl(fen_ctx* ctx, int vector_in_x, int vector_out_x);                      /* fields.f90:347 */

/* ---- poisson_mod (src/poisson.f90:51-52) --------------------------------------------------- */
int fen_gpu_init_poisson_solver(fen_ctx* ctx);                /* poisson.f90:57   */
int fen_gpu_solve_poisson(fen_ctx* ctx, int field);           /* solve_poisson pointer, :29 */
int fen_gpu_destroy_poisson_solver(fen_ctx* ctx);             /* poisson.f90:1456 */
/* variant chosen by init: "pp","pn","nn","ppp","ppn","npn","nnn" */
const char* fen_gpu_poisson_variant(fen_ctx* ctx);

/* ---- solver_mod / navier_stokes_mod -------------------------------------------------------- */
/* init_solver (solver.f90:34): allocate_navier_stokes_fields (navier_stokes.f90:752: fields,
 * rho = density, mu = viscosity, BC wiring table) + init_poisson_solver */
int fen_gpu_init_solver(fen_ctx* ctx);
int fen_gpu_destroy_solver(fen_ctx* ctx);                     /* solver.f90:333 */

/* module scalars of navier_stokes_mod (navier_stokes.f90:18-45) */
typedef struct fen_ns_params {
    double density, viscosity;
    double g[3];
    double CFL;
    double dt_o;
    double dt_visc, dt_conv;
    int constant_CFL;
} fen_ns_params;
int fen_gpu_get_params(fen_ctx* ctx, fen_ns_params* p);
int fen_gpu_set_params(fen_ctx* ctx, const fen_ns_params* p);

/* set_timestep(comp_grid, dt, U) (navier_stokes.f90:623): also sets dt_o = dt */
int fen_gpu_set_timestep(fen_ctx* ctx, double U, double* dt);
/* navier_stokes_solver(comp_grid, step, dt) == advance_solution (navier_stokes.f90:50,
 * solver.f90:75).  dt is inout (changes only with constant_CFL).  Asynchronous unless
 * constant_CFL: maxdiv / maxCFL are fetched by fen_gpu_get_status. */
int fen_gpu_navier_stokes_solver(fen_ctx* ctx, int step, double* dt);
/* maxdiv, maxCFL of the last step (navier_stokes.f90:24,593,617); synchronises */
int fen_gpu_get_status(fen_ctx* ctx, double* maxdiv, double* maxCFL);
/* print_navier_stokes_solver_status (navier_stokes.f90:734): formats the line into buf */
int fen_gpu_status_line(fen_ctx* ctx, int step, double time, double dt, char* buf, int buflen);

/* stage-level entry points the reference's tests call directly */
int fen_gpu_add_advection(fen_ctx* ctx, int rhs_vector_x);             /* navier_stokes.f90:261 */
int fen_gpu_compute_explicit_terms(fen_ctx* ctx, int rhs_vector_x);    /* navier_stokes.f90:217 */
int fen_gpu_predicted_velocity_field(fen_ctx* ctx, double dt);         /* navier_stokes.f90:140 */
int fen_gpu_correct_velocity_field(fen_ctx* ctx, double dt);           /* navier_stokes.f90:505 */
int fen_gpu_update_pressure(fen_ctx* ctx);                             /* navier_stokes.f90:550 */
int fen_gpu_checks(fen_ctx* ctx, double dt);                           /* navier_stokes.f90:570 */

/* ---- raw field files in the reference's on-disk format (global interior array, x fastest, real(dp), no header;
 * 2decomp MPI-IO).  Collective over the ranks; every rank reads / writes its own z-slab byte range.  As with the
 * reference's MPI-IO calls the caller synchronises around them: a file written by save_* is complete for the OTHER
 * ranks only after a caller-side barrier, and the first I/O call of a context allocates its (kept) staging buffers,
 * which synchronises the device -- make it between steps, when no rank is inside a collective. ------------------ */
int fen_gpu_scalar_write(fen_ctx* ctx, int field, const char* filename);   /* scalar%write, scalar.f90:428 */
int fen_gpu_scalar_read(fen_ctx* ctx, int field, const char* filename);    /* scalar%read,  scalar.f90:400 */
/* save_state / load_state (solver.f90:160, :244): p, v_x, v_y, dv_o_x, dv_o_y, [v_z, dv_o_z] in one file; load also
 * updates the ghost nodes of p and v.  dt and dt_o are not stored (the driver re-derives them, test_NS.f90:54-63). */
int fen_gpu_save_state(fen_ctx* ctx, const char* filename);
int fen_gpu_load_state(fen_ctx* ctx, const char* filename);
/* save_fields(step) (solver.f90:103): <dir>/vx_<step>.raw, vy_, [vz_] (cell-centred, fields.f90:210) and p_ */
int fen_gpu_save_fields(fen_ctx* ctx, int step, const char* dir);

/* ---- host hook between the predictor and the Poisson right-hand side: the place of apply_ibm_forcing(v, dt)
 * (navier_stokes.f90:106-108).  The callback may pull v, force it on the host and push it back; the library then
 * refreshes the ghost nodes of v.  A non-zero return aborts the step.  NULL removes the hook. ------------------- */
typedef int (*fen_forcing_fn)(void* user, int step, double dt);
int fen_gpu_set_forcing_hook(fen_ctx* ctx, fen_forcing_fn fn, void* user);

/* ---- two-phase path: volume_of_fluid_mod (MTHINC VoF, src/volume_of_fluid.f90), multiphase_mod
 * (src/multiphase.f90) and the `#ifdef MF` branches of navier_stokes_mod / solver_mod.  As in the reference this
 * path is 2-D only (ndim = 2): compute_norm, compute_flux and the variable-viscosity stress divergence have no z
 * terms (volume_of_fluid.f90:307-396, 558-642; navier_stokes.f90:405-452).  The reference selects it at compile time
 * (-DMF); here fen_gpu_init_solver_mf selects it per context, after which fen_gpu_set_timestep and
 * fen_gpu_navier_stokes_solver take the MF branches. ------------------------------------------------------------- */
typedef struct fen_mf_params {
    double rho_0, rho_1, mu_0, mu_1;   /* multiphase.f90:18 */
    double sigma;                      /* multiphase.f90:21 */
    double beta;                       /* volume_of_fluid.f90:24 sharpness */
    double cut;                        /* volume_of_fluid.f90:37 */
    int quadratic;                     /* volume_of_fluid.f90:27 */
    int x_first;                       /* volume_of_fluid.f90:30 (toggled by every advect_vof) */
    double dt_surf;                    /* navier_stokes.f90:29; set by set_timestep when sigma > 0, +inf before */
    double rhomin, irhomin;            /* multiphase.f90:29; set by init_solver_mf (solver.f90:92-93) */
} fen_mf_params;
int fen_gpu_mf_get_params(fen_ctx* ctx, fen_mf_params* p);
int fen_gpu_mf_set_params(fen_ctx* ctx, const fen_mf_params* p);
/* allocate_vof_fields (volume_of_fluid.f90:54): vof, h, d, curv, norm, l with one ghost layer, Periodic -> 0 and
 * Wall -> Neumann on the four faces.  Usable without init_solver (the reference's VoF-only tests do that). */
int fen_gpu_allocate_vof_fields(fen_ctx* ctx);
/* get_vof_from_distance (volume_of_fluid.f90:676): `fn` is the reference's `distance` procedure pointer (:40-46).  It
 * is evaluated on the host at the cell centre and the four Gauss points of every cell; the tanh profile, the
 * quadrature and the ghost update run on the device.  NULL -> FEN_ERR_ARG ('ERROR: distance function not defined.') */
typedef double (*fen_distance_fn)(void* user, double x, double y);
int fen_gpu_get_vof_from_distance(fen_ctx* ctx, fen_distance_fn fn, void* user, double x0, double y0);
int fen_gpu_get_h_from_vof(fen_ctx* ctx);                                   /* volume_of_fluid.f90:228 */
int fen_gpu_advect_vof(fen_ctx* ctx, int vector_x, double dt);              /* volume_of_fluid.f90:434 */
int fen_gpu_check_vof_integral(fen_ctx* ctx, double* int_phase_1, double* int_phase_2);   /* :722 */
int fen_gpu_destroy_vof(fen_ctx* ctx);                                      /* volume_of_fluid.f90:758 */
int fen_gpu_update_material_properties(fen_ctx* ctx);                       /* multiphase.f90:121 */
/* init_solver compiled with -DMF (solver.f90:34-99): fen_gpu_init_solver + constant_viscosity = .false. +
 * allocate_vof_fields + allocate_multiphase_fields (multiphase.f90:46) + get_vof_from_distance (skipped with the
 * reference's error message when fn is NULL: push FEN_VOF and call fen_gpu_update_material_properties instead) +
 * update_material_properties + rhomin.  (x0, y0) = grid origin (grid%x(i) = x0 + (i - 1/2) delta, grid.f90:155). */
int fen_gpu_init_solver_mf(fen_ctx* ctx, fen_distance_fn fn, void* user, double x0, double y0);

/* ---- measurement helpers (no reference analogue: FEN has no timers) --- */
/* per-kernel-family CUDA-event timing of the next steps; names/ms arrays sized by the caller */
int fen_gpu_profile_enable(fen_ctx* ctx, int on);
int fen_gpu_profile_read(fen_ctx* ctx, int max_entries, char names[][32], double* ms, int* launches,
                         int* n_out);
/* total kernel launches issued by this context since creation */
long long fen_gpu_launch_count(fen_ctx* ctx);
/* CUDA stream of the context (as void*), for callers that time with their own events */
void* fen_gpu_stream(fen_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* FEN_GPU_H */
