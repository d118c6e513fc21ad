// Host loops over the __host__ __device__ cell functions of fen_b200/csrc/vof_math.cuh -- the same source the CUDA
// kernels of multiphase.cu compile -- so that the transcription of the reference's VoF / two-phase arithmetic can be
// held against the oracle without a GPU.  Built by tests/test_vof_math_host.py with g++ (once without and once with
// FMA contraction, which is what nvcc does on the device).  TEST CODE: not part of libfen_gpu.so.
//
// Arrays are the oracle's 2-D slices: Fortran order, one ghost layer, element (i, j) at (i) + (nx + 2) * (j) with
// i in [0, nx+1], j in [0, ny+1]; outputs are interior-only, (i - 1) + nx * (j - 1).
#include "../../fen_b200/csrc/vof_math.cuh"

using namespace fen;

#define G(a, i, j) (a)[(size_t)(i) + (size_t)(nx + 2) * (size_t)(j)]
#define I(a, i, j) (a)[(size_t)((i) - 1) + (size_t)nx * (size_t)((j) - 1)]

extern "C" {

void host_recon(int nx, int ny, const double* vof, double delta, double beta, double cut, int quadratic, double* onx,
                double* ony, double* olx, double* oly, double* ocurv, double* oh, double* od) {
    const double id = 1.0 / delta, id2 = 1.0 / (delta * delta);
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            double s[3][3];
            for (int b = 0; b < 3; ++b)
                for (int a = 0; a < 3; ++a) s[b][a] = G(vof, i + a - 1, j + b - 1);
            const VofRecon r = vof_norm(s, delta, id, id2, quadratic != 0);
            double h, d;
            vof_h_d(s[1][1], r.nx, r.ny, r.lx, r.ly, beta, cut, h, d);
            I(onx, i, j) = r.nx; I(ony, i, j) = r.ny; I(olx, i, j) = r.lx; I(oly, i, j) = r.ly;
            I(ocurv, i, j) = r.curv; I(oh, i, j) = h; I(od, i, j) = d;
        }
}

// the tiled reconstruction (k_vof_recon_tile): blocks and threads are loops, the barriers are the phase boundaries
void host_recon_tile(int nx, int ny, const double* vof, double delta, double beta, double cut, int quadratic,
                     double* onx, double* ony, double* olx, double* oly, double* ocurv, double* oh, double* od) {
    const double id = 1.0 / delta, id2 = 1.0 / (delta * delta);
    const long long sy = nx + 2;
    static VofTile T;
    for (int by = 0; by < (ny + VT_Y - 1) / VT_Y; ++by)
        for (int bx = 0; bx < (nx + VT_X - 1) / VT_X; ++bx) {
            const int i0 = bx * VT_X, j0 = by * VT_Y;
            const int wa = VT_FW < nx + 2 - i0 ? VT_FW : nx + 2 - i0, hb = VT_FH < ny + 2 - j0 ? VT_FH : ny + 2 - j0;
            for (int tid = 0; tid < VT_N; ++tid) vof_tile_load(T, tid, vof + i0 + sy * j0, sy, wa, hb);
            for (int tid = 0; tid < VT_N; ++tid) vof_tile_corners(T, tid, id);
            for (int ty = 0; ty < VT_Y; ++ty)
                for (int tx = 0; tx < VT_X; ++tx) {
                    const int i = i0 + 1 + tx, j = j0 + 1 + ty;
                    if (i > nx || j > ny) continue;
                    double v00, h, d;
                    const VofRecon r = vof_tile_cell(T, tx, ty, delta, id2, quadratic != 0, v00);
                    vof_h_d(v00, r.nx, r.ny, r.lx, r.ly, beta, cut, h, d);
                    I(onx, i, j) = r.nx; I(ony, i, j) = r.ny; I(olx, i, j) = r.lx; I(oly, i, j) = r.ly;
                    I(ocurv, i, j) = r.curv; I(oh, i, j) = h; I(od, i, j) = d;
                }
        }
}

void host_sweep(int nx, int ny, int dir, int final_, int x_first, const double* src, const double* fnx,
                const double* fny, const double* flx, const double* fly, const double* fd, const double* u,
                const double* v, double dt, double delta, double beta, double cut, double* out) {
    const int si = dir == 1 ? 1 : 0, sj = dir == 1 ? 0 : 1;
    const double* vel = dir == 1 ? u : v;
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            const double up = G(vel, i, j), um = G(vel, i - si, j - sj);
            const int ipc = up >= 0.0 ? i : i + si, jpc = up >= 0.0 ? j : j + sj;
            const int imc = um >= 0.0 ? i - si : i, jmc = um >= 0.0 ? j - sj : j;
            const double fp = vof_flux(dir, up, dt, delta, beta, cut, G(src, ipc, jpc), G(fnx, ipc, jpc),
                                       G(fny, ipc, jpc), G(flx, ipc, jpc), G(fly, ipc, jpc), G(fd, ipc, jpc));
            const double fm = vof_flux(dir, um, dt, delta, beta, cut, G(src, imc, jmc), G(fnx, imc, jmc),
                                       G(fny, imc, jmc), G(flx, imc, jmc), G(fly, imc, jmc), G(fd, imc, jmc));
            const double s0 = G(src, i, j);
            const double val = (s0 - (fp - fm) / delta) / (1.0 - dt * (up - um) / delta);
            if (!final_) { I(out, i, j) = val; continue; }
            const double dux = G(u, i, j) - G(u, i - 1, j);
            const double dvy = G(v, i, j) - G(v, i, j - 1);
            if (x_first) I(out, i, j) = val - dt * (s0 * dux / delta + val * dvy / delta);
            else I(out, i, j) = val - dt * (val * dux / delta + s0 * dvy / delta);
        }
}

void host_predict(int nx, int ny, const double* u, const double* v, const double* p, const double* ph,
                  const double* rho, const double* mu, const double* vof, const double* curv, const double* sx,
                  const double* sy, const double* dvox, const double* dvoy, double id, double dt, double A, double B,
                  double g0, double g1, double sigma, double irhomin, double* un, double* vn, double* odvx,
                  double* odvy) {
    MfPrm k{id, dt, A, B, g0, g1, sigma, irhomin, 1};
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            MfCell q;
            for (int b = 0; b < 3; ++b)
                for (int a = 0; a < 3; ++a) {
                    q.u[b][a] = G(u, i + a - 1, j + b - 1);
                    q.v[b][a] = G(v, i + a - 1, j + b - 1);
                    q.m[b][a] = G(mu, i + a - 1, j + b - 1);
                }
            q.rho0 = G(rho, i, j); q.rhoip = G(rho, i + 1, j); q.rhojp = G(rho, i, j + 1);
            q.p0 = G(p, i, j); q.pip = G(p, i + 1, j); q.pjp = G(p, i, j + 1);
            q.h0 = G(ph, i, j); q.hip = G(ph, i + 1, j); q.hjp = G(ph, i, j + 1);
            q.c0 = G(curv, i, j); q.cip = G(curv, i + 1, j); q.cjp = G(curv, i, j + 1);
            q.f0 = G(vof, i, j); q.fip = G(vof, i + 1, j); q.fjp = G(vof, i, j + 1);
            q.sx = I(sx, i, j); q.sy = I(sy, i, j);
            q.dvox = I(dvox, i, j); q.dvoy = I(dvoy, i, j);
            double a_, b_, c_, d_;
            mf_predict_cell(q, k, a_, b_, c_, d_);
            I(un, i, j) = a_; I(vn, i, j) = b_; I(odvx, i, j) = c_; I(odvy, i, j) = d_;
        }
}

}  // extern "C"
