// tgv_driver.cpp -- the reference's 2-D Taylor-Green driver (test/small_test/navier_stokes/taylor_green_vortex/
// taylor_green_vortex.f90) written against include/fen_gpu.hpp, the C++ mirror of FEN's solver API: grid%setup,
// init_solver, set_timestep, the initial condition written into the host arrays of v and p, the time loop with
// advance_solution and print_solver_status, and the two explicit transfer points (push / pull).
//
//   g++ -std=c++17 -Iinclude examples/tgv_driver.cpp -Lfen_b200 -lfen_gpu -Wl,-rpath,$PWD/fen_b200 -o tgv_driver_cpp
//   ./tgv_driver_cpp 64 40      # N, steps -- needs a CUDA device; without one grid%setup throws (no CPU fallback)
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "fen_gpu.hpp"

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 64;
    const int nstep = argc > 2 ? std::atoi(argv[2]) : 10;
    const double L = 2.0 * std::acos(-1.0);
    try {
        fen::grid comp_grid;
        comp_grid.setup(n, n, 1, L, L, L / n);                       // :36-43 (one rank: prow = pcol = 1)
        fen::solver ns(comp_grid);
        ns.density = 1.0;
        ns.viscosity = 1.0;
        ns.init_solver();
        double dt = ns.set_timestep(2.0);
        const double d = comp_grid.delta;
        for (int j = 1; j <= n; ++j)                                // :86-113, staggered locations
            for (int i = 1; i <= n; ++i) {
                const double xc = comp_grid.x(i), yc = comp_grid.y(j), xf = i * d, yf = j * d;
                ns.v.x(i, j) = -std::cos(xf) * std::sin(yc);
                ns.v.y(i, j) = std::sin(xc) * std::cos(yf);
                ns.p(i, j) = -0.25 * (std::cos(2.0 * xc) + std::cos(2.0 * yc));
            }
        ns.v.push();
        ns.p.push();
        ns.v.update_ghost_nodes();
        ns.p.update_ghost_nodes();
        double time = 0.0;
        for (int step = 1; step <= nstep; ++step) {
            time += dt;
            ns.advance_solution(step, dt);
            const std::string line = ns.print_solver_status(step, time, dt);
            if (step == 1 || step == nstep) std::printf("%s\n", line.c_str());
        }
        ns.v.pull();
        double emax = 0.0;                                          // u = -cos x sin y exp(-2 nu t) (postpro.py:48)
        for (int j = 1; j <= n; ++j)
            for (int i = 1; i <= n; ++i)
                emax = std::fmax(emax, std::fabs(ns.v.x(i, j) + std::cos(i * d) * std::sin(comp_grid.y(j)) * std::exp(-2.0 * time)));
        std::printf("max |u - u_exact| after %d steps: %.3e\n", nstep, emax);
        std::printf("poisson variant %s\n", ns.poisson_variant().c_str());
        ns.destroy_solver();
        comp_grid.destroy();
    } catch (const fen::error& e) {
        std::fprintf(stderr, "fen error %d: %s\n", e.code, e.what());
        return e.code;
    }
    return 0;
}
