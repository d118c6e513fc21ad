// ghost.cu -- scalar%update_ghost_nodes / vector%update_ghost_nodes on the device.
//
// Restates src/scalar.f90:223-396: halo exchange (src/halo.f90:12-43, z neighbours only because the
// decomposition is (prow, pcol) = (1, nranks)), then the physical boundary conditions in the
// reference's order x-left, x-right, y-bottom, y-top, z-front, z-back, each over the FULL
// transverse extent (ghost rows/planes included) -- that order defines the edge and corner ghosts
// the advection stencil reads (SURVEY.md hazard H3).  Up to eight fields are handled per launch (the
// components of a vector, or the seven fields one VoF reconstruction refreshes).
#include "fen_internal.cuh"

namespace fen {

struct GhostField {
    double* f;
    int tlo, thi;          // bc type on the low / high face of this direction
    int normal;            // 1 if this component is the wall-normal staggered one for this direction
    int mlo, mhi;          // BcMode
    double vlo, vhi;       // uniform value
    const double* plo;     // value plane (incl. ghosts), Fortran order
    const double* phi;
};
constexpr int GHOST_MAX = 8;
struct GhostArgs {
    GhostField fld[GHOST_MAX];
    int n;
    Layout L;
};

__device__ __forceinline__ double bc_val(int mode, double v, const double* plane, long long at) {
    return mode == BC_ZERO ? 0.0 : (mode == BC_UNIFORM ? v : plane[at]);
}

// DIR 0: x faces, threads over (j, k); DIR 1: y faces, threads over (i, k); DIR 2: z faces, (i, j).
template <int DIR>
__global__ void k_ghost(GhostArgs a) {
    const Layout& L = a.L;
    const int n0 = (DIR == 0) ? L.ny + 2 : L.nx + 2;
    const int n1 = (DIR == 2) ? L.ny + 2 : L.nzl + 2;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (p >= n0 || q >= n1) return;
    const int n = (DIR == 0) ? L.nx : (DIR == 1 ? L.ny : L.nzl);
    const long long stride = (DIR == 0) ? 1 : (DIR == 1 ? L.sy : L.sz);
    long long base;      // element with index 0 along DIR
    if (DIR == 0) base = L.idx(0, p, q);
    else if (DIR == 1) base = L.idx(p, 0, q);
    else base = L.idx(p, q, 0);
    const long long at = (long long)p + (long long)n0 * q;   // position inside the bc value plane
    for (int m = 0; m < a.n; ++m) {
        const GhostField& g = a.fld[m];
        double* f = g.f + base;
        // low face (scalar.f90:255-271, 294-316, 345-365)
        if (g.tlo == FEN_PERIODIC) {
            f[0] = f[n * stride];
        } else if (g.tlo == FEN_DIRICHLET) {
            double b = bc_val(g.mlo, g.vlo, g.plo, at);
            f[0] = g.normal ? b : 2.0 * b - f[stride];
        } else if (g.tlo == FEN_NEUMANN) {
            f[0] = f[stride];
        }
        // high face (scalar.f90:274-291, 319-342, 367-388)
        if (g.thi == FEN_PERIODIC) {
            f[(n + 1) * stride] = f[stride];
        } else if (g.thi == FEN_DIRICHLET) {
            double b = bc_val(g.mhi, g.vhi, g.phi, at);
            if (g.normal) {
                f[n * stride] = b;
                f[(n + 1) * stride] = b;
            } else {
                f[(n + 1) * stride] = 2.0 * b - f[n * stride];
            }
        } else if (g.thi == FEN_NEUMANN) {
            f[(n + 1) * stride] = f[n * stride];
        }
    }
}

int ghost_update(fen_ctx* c, int field, int ncomp, bool x_done) {
    if (ncomp < 1 || ncomp > 3) return set_error(FEN_ERR_ARG, "update_ghost_nodes: ncomp must be 1..3");
    const int ids[3] = {field, field + 1, field + 2};
    return ghost_update_list(c, ids, ncomp, x_done);
}

int ghost_update_list(fen_ctx* c, const int* ids, int ncomp, bool x_done) {
    if (ncomp < 1 || ncomp > GHOST_MAX) return set_error(FEN_ERR_ARG, "update_ghost_nodes: 1..%d fields per call", GHOST_MAX);
    Field* fp[GHOST_MAX];
    double* ptrs[GHOST_MAX];
    for (int m = 0; m < ncomp; ++m) {
        FEN_TRY(field_check(c, ids[m], &fp[m]));
        if (fp[m]->gl < 1)   // vector.f90:91-95 prints an error and skips
            return set_error(FEN_ERR_ARG, "Cannot update ghost nodes on a scalar without ghost nodes (field %d)",
                             ids[m]);
        ptrs[m] = fp[m]->d;
    }
    // 1. halo exchange with the z neighbours (scalar.f90:251 -> halo.f90:33), three fields per round
    if (c->g.nranks > 1)
        for (int m0 = 0; m0 < ncomp; m0 += 3) FEN_TRY(halo_exchange(c, ptrs + m0, std::min(3, ncomp - m0)));

    const Layout& L = c->L;
    const int ndir = (c->g.ndim == 3) ? 3 : 2;
    for (int dir = 0; dir < ndir; ++dir) {
        GhostArgs a;
        a.n = ncomp;
        a.L = L;
        bool any = false;
        for (int m = 0; m < ncomp; ++m) {
            Field& F = *fp[m];
            GhostField& g = a.fld[m];
            g.f = F.d;
            const int flo = 2 * dir, fhi = 2 * dir + 1;
            g.tlo = F.bc_type[flo];
            g.thi = F.bc_type[fhi];
            // periodic z ghosts come from the halo exchange when the slab is split
            // (scalar.f90:348 `if (self%G%pcol == 1)`)
            if (dir == 2 && c->g.nranks > 1) {
                if (g.tlo == FEN_PERIODIC) g.tlo = FEN_HALO;
                if (g.thi == FEN_PERIODIC) g.thi = FEN_HALO;
            }
            g.normal = (F.loc == FEN_LOC_X + dir) ? 1 : 0;
            g.mlo = F.bc_mode[flo];
            g.mhi = F.bc_mode[fhi];
            g.vlo = F.bc_value[flo];
            g.vhi = F.bc_value[fhi];
            g.plo = F.bc_plane[flo];
            g.phi = F.bc_plane[fhi];
            for (int t : {g.tlo, g.thi}) {
                if (t < -1 || t > 2)
                    return set_error(FEN_ERR_ARG, "wrong boundary condition type %d for field %d", t, ids[m]);
                if (t != FEN_HALO) any = true;
            }
        }
        if (!any) continue;
        if (dir == 0 && x_done) {
            // the producing kernel has already written the periodic x ghosts of every interior row; the y and z
            // passes below copy whole rows / planes, so the edge and corner ghosts still come out in the
            // reference's order (scalar.f90:255-388)
            bool per = true;
            for (int m = 0; m < ncomp; ++m) per = per && a.fld[m].tlo == FEN_PERIODIC && a.fld[m].thi == FEN_PERIODIC;
            if (per) continue;
        }
        const int n0 = (dir == 0) ? L.ny + 2 : L.nx + 2;
        const int n1 = (dir == 2) ? L.ny + 2 : L.nzl + 2;
        dim3 block(128), grid((n0 + 127) / 128, n1);
        if (dir == 0) FEN_LAUNCH(c, "ghost_x", k_ghost<0><<<grid, block, 0, c->stream>>>(a));
        if (dir == 1) FEN_LAUNCH(c, "ghost_y", k_ghost<1><<<grid, block, 0, c->stream>>>(a));
        if (dir == 2) FEN_LAUNCH(c, "ghost_z", k_ghost<2><<<grid, block, 0, c->stream>>>(a));
    }
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

}  // namespace fen
